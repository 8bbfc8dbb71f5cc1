// pool.cu -- what happens to the sample set and to the minimiser pool around the star test
// (rows a4' and f3).  Small, HBM / latency bound integer and compare work.
//
//   subspace_mask   SHGO.sampling_subspace (reference _shgo.py:1452-1463): keep the sample points with
//                   every g_j(x) >= 0 BEFORE triangulation; a NaN constraint value drops the point
//                   (`nan >= 0.0` is False) -- the opposite of feasibility_check's rule.
//   gather_rows     the surviving points as the C-order (m, d) array the driver hands to qhull.
//   pool_rank       SHGO.minimizers' post-processing (_shgo.py:1116-1122, 1213-1219) and g_topograph
//                   (_shgo.py:1227-1243): drop the pool vertices whose coordinates equal an earlier
//                   local minimum, order the rest by f (or by distance from a point), so that only a
//                   ranked short list returns to the host.
#include "common.cuh"

namespace shgo_b200 {

static inline size_t align256p(size_t x) { return (x + 255) & ~(size_t)255; }
static inline unsigned grid1p(uint64_t items, int per_sm)
{
    uint64_t b = (items + 255) / 256;
    const uint64_t cap = (uint64_t)sm_count() * per_sm;
    if (b > cap) b = cap;
    return (unsigned)(b < 1 ? 1 : b);
}

__global__ void __launch_bounds__(256)
subspace_mask_kernel(const double *__restrict__ g, int m, uint64_t n, uint8_t *__restrict__ keep)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        bool ok = true;
        for (int c = 0; c < m; ++c) ok = ok && (ld_nc_f64(g + (uint64_t)c * n + i) >= 0.0);  // NaN: dropped
        keep[i] = ok ? 1 : 0;
    }
}

__global__ void __launch_bounds__(256)
gather_rows_kernel(const double *__restrict__ x_soa, uint64_t ld, int d, const int64_t *__restrict__ idx,
                   uint64_t m, double *__restrict__ out)
{
    const uint64_t total = m * (uint64_t)d;
    for (uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; e < total;
         e += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = e / d;
        const int j = (int)(e - r * d);
        out[e] = ld_nc_f64(x_soa + (uint64_t)j * ld + idx[r]);
    }
}

// order-preserving map double -> uint64 (-inf < ... < -0 = +0 < ... < +inf < NaN-with-sign-clear)
__device__ __forceinline__ unsigned long long orderable(double v)
{
    if (v == 0.0) v = 0.0;  // -0.0 and +0.0 are one key (they compare equal in the reference's argsort)
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// one thread per pool entry: excluded (coordinates equal to one of the k earlier minima) entries get
// the largest key; mode 0: key = f, mode 1: key = euclidean distance from xref (the sum of squares is
// accumulated axis by axis like scipy's cdist, so that the order agrees with the reference's argsort)
__global__ void __launch_bounds__(256)
pool_key_kernel(const int64_t *__restrict__ pool, uint64_t m, const double *__restrict__ f,
                const double *__restrict__ x_soa, uint64_t ld, int d, const double *__restrict__ xl, int k,
                int mode, const double *__restrict__ xref, uint32_t *__restrict__ key_lo,
                uint32_t *__restrict__ key_hi, int32_t *__restrict__ ids, unsigned long long *__restrict__ nkept)
{
    unsigned kept = 0;
    for (uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; p < m;
         p += (uint64_t)gridDim.x * blockDim.x) {
        const int64_t v = pool[p];
        bool excluded = false;
        for (int q = 0; q < k && !excluded; ++q) {
            bool same = true;
            for (int j = 0; j < d && same; ++j) same = x_soa[(uint64_t)j * ld + v] == xl[q * d + j];
            excluded = same;
        }
        unsigned long long key = ~0ull;
        if (!excluded) {
            double val;
            if (mode == 0) {
                val = f[v];
            } else {
                double s = 0.0;
                for (int j = 0; j < d; ++j) {
                    const double df = __dsub_rn(x_soa[(uint64_t)j * ld + v], xref[j]);
                    s = __dadd_rn(s, __dmul_rn(df, df));
                }
                val = __dsqrt_rn(s);
            }
            key = orderable(val);
            if (key == ~0ull) key = ~0ull - 1;  // keep the all-ones key for the excluded entries
            ++kept;
        }
        key_lo[p] = (uint32_t)key;
        key_hi[p] = (uint32_t)(key >> 32);
        ids[p] = (int32_t)p;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kept += __shfl_down_sync(0xffffffffu, kept, o);
    if ((threadIdx.x & 31) == 0 && kept) atomicAdd(nkept, (unsigned long long)kept);
}

__global__ void __launch_bounds__(256)
permute_u32_kernel(const uint32_t *__restrict__ src, const int32_t *__restrict__ ids, uint64_t m,
                   uint32_t *__restrict__ dst)
{
    for (uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; p < m;
         p += (uint64_t)gridDim.x * blockDim.x)
        dst[p] = src[ids[p]];
}

__global__ void __launch_bounds__(256)
pool_emit_kernel(const int64_t *__restrict__ pool, const int32_t *__restrict__ ids, const uint64_t *__restrict__ nkept,
                 int64_t *__restrict__ ranked)
{
    const uint64_t m = *nkept;
    for (uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; p < m;
         p += (uint64_t)gridDim.x * blockDim.x)
        ranked[p] = pool[ids[p]];
}

}  // namespace shgo_b200

using namespace shgo_b200;

extern "C" {

int shgo_b200_subspace_mask(const double *g, int m, uint64_t n, uint8_t *keep, void *stream)
{
    if (n == 0) return SHGO_B200_OK;
    SHGO_REQUIRE(keep, "null output");
    SHGO_REQUIRE(m == 0 || g, "m > 0 but g is null");
    SHGO_REQUIRE(m >= 0, "negative constraint count");
    subspace_mask_kernel<<<grid1p(n, 8), 256, 0, as_stream(stream)>>>(g, m, n, keep);
    SHGO_CUDA_TRY(cudaGetLastError());
    return SHGO_B200_OK;
}

int shgo_b200_gather_rows(const double *x_soa, uint64_t ld, int d, const int64_t *idx, uint64_t m, double *out,
                          void *stream)
{
    SHGO_REQUIRE(d >= 1, "d must be positive (got %d)", d);
    if (m == 0) return SHGO_B200_OK;
    SHGO_REQUIRE(x_soa && idx && out, "null pointer");
    gather_rows_kernel<<<grid1p(m * (uint64_t)d, 8), 256, 0, as_stream(stream)>>>(x_soa, ld, d, idx, m, out);
    SHGO_CUDA_TRY(cudaGetLastError());
    return SHGO_B200_OK;
}

size_t shgo_b200_pool_rank_workspace(uint64_t m)
{
    // key lo / hi / scratch keys, ids x2, histogram, counter
    return 3 * align256p(m * 4) + 2 * align256p(m * 4) + radix_hist_bytes(m) + align256p(16);
}

int shgo_b200_pool_rank(const int64_t *pool, uint64_t m, const double *f, const double *x_soa, uint64_t ld,
                        int d, const double *xl, int k, int mode, const double *xref, int64_t *ranked,
                        uint64_t *count, void *tmp, size_t tmp_bytes, void *stream)
{
    SHGO_REQUIRE(count, "null count");
    SHGO_REQUIRE(d >= 1 && k >= 0 && (mode == 0 || mode == 1), "bad arguments");
    SHGO_REQUIRE(m <= 0x7fffffffull, "pool too large");
    cudaStream_t st = as_stream(stream);
    SHGO_CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(uint64_t), st));
    if (m == 0) return SHGO_B200_OK;
    SHGO_REQUIRE(pool && ranked && tmp, "null pointer");
    SHGO_REQUIRE(mode == 1 || f, "null f");
    SHGO_REQUIRE((k == 0 && mode == 0) || x_soa, "coordinates are needed to exclude minima / rank by distance");
    SHGO_REQUIRE(k == 0 || xl, "null xl");
    SHGO_REQUIRE(mode == 0 || xref, "null xref");
    if (tmp_bytes < shgo_b200_pool_rank_workspace(m))
        return fail(SHGO_B200_ERR_WORKSPACE, "pool-rank workspace: need %zu bytes, got %zu",
                    shgo_b200_pool_rank_workspace(m), tmp_bytes);
    unsigned char *p = static_cast<unsigned char *>(tmp);
    uint32_t *klo = reinterpret_cast<uint32_t *>(p);
    p += align256p(m * 4);
    uint32_t *khi = reinterpret_cast<uint32_t *>(p);
    p += align256p(m * 4);
    uint32_t *k1 = reinterpret_cast<uint32_t *>(p);
    p += align256p(m * 4);
    int32_t *i0 = reinterpret_cast<int32_t *>(p);
    p += align256p(m * 4);
    int32_t *i1 = reinterpret_cast<int32_t *>(p);
    p += align256p(m * 4);
    uint64_t *hist = reinterpret_cast<uint64_t *>(p);
    pool_key_kernel<<<grid1p(m, 8), 256, 0, st>>>(pool, m, f, x_soa, ld, d, xl, k, mode, xref, klo, khi, i0,
                                                 reinterpret_cast<unsigned long long *>(count));
    // 64-bit keys: stable sort by the low word, then by the high word of the permuted keys
    if (int rc = radix_sort_pairs(klo, i0, k1, i1, hist, m, st)) return rc;
    permute_u32_kernel<<<grid1p(m, 8), 256, 0, st>>>(khi, i0, m, klo);
    if (int rc = radix_sort_pairs(klo, i0, k1, i1, hist, m, st)) return rc;
    pool_emit_kernel<<<grid1p(m, 8), 256, 0, st>>>(pool, i0, count, ranked);
    SHGO_CUDA_TRY(cudaGetLastError());
    return SHGO_B200_OK;
}

}  // extern "C"
