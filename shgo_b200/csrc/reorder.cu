// reorder.cu -- locality-preserving vertex numbering for explicit meshes (SURVEY.md section 8(e):
// "renumber vertices (e.g. Morton order of the points) so that most neighbours are local").
//
// qhull numbers the vertices of a Delaunay mesh like the points it was given; Sobol order carries
// no spatial locality, so the star test's neighbour gathers of f are random 8-byte reads (round 1:
// 0.117 of the HBM roofline on such a graph).  Sorting the vertices along a Z-order curve makes the
// neighbours of consecutive rows land in the same few cache lines.  This file provides
//   morton_order   32-bit Z-order key of every point (floor(32/d) bits per axis) + stable LSD radix
//                  sort (4 passes of 8 bits) -> order[new] = old, newid[old] = new
//   relabel3       simplices [S][dp1] -> first three slots, renumbered ([S][3]: all vf_to_vv reads,
//                  reference _complex.py:1064-1069)
//   gather / scatter   f and flags between the two numberings
// The classification result is the same set of vertices in either numbering (tests compare them).
// Integer / byte work bound by HBM; no tensor cores involved.
#include "common.cuh"

namespace shgo_b200 {

static inline size_t align256r(size_t x) { return (x + 255) & ~(size_t)255; }

// ---- Z-order keys -------------------------------------------------------------------------------------
// axis j contributes bit t (t = 0 most significant) at key bit  (bits-1-t)*dk + (dk-1-j),  dk = min(d, 32)
__global__ void __launch_bounds__(256)
morton_key_kernel(const double *__restrict__ x_soa, uint64_t ld, int d, uint64_t n,
                  const double *__restrict__ lb, const double *__restrict__ ub, uint32_t *__restrict__ keys,
                  int32_t *__restrict__ ids)
{
    const int dk = d < 32 ? d : 32;
    const int bits = dk == 1 ? 30 : 32 / dk;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t key = 0;
        for (int j = 0; j < dk; ++j) {
            const double lo = __ldg(lb + j), hi = __ldg(ub + j);
            const double w = hi - lo;
            double t = w > 0.0 ? (ld_nc_f64(x_soa + (uint64_t)j * ld + i) - lo) / w : 0.0;
            t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);  // NaN -> 0 through the first comparison being false
            if (!(t == t)) t = 0.0;
            uint32_t q = (uint32_t)(t * (double)(1u << bits));
            if (q >= (1u << bits)) q = (1u << bits) - 1u;
            for (int b = 0; b < bits; ++b)
                key |= ((q >> b) & 1u) << (b * dk + (dk - 1 - j));
        }
        keys[i] = key;
        ids[i] = (int32_t)i;
    }
}

// ---- stable LSD radix sort, 8 bits per pass --------------------------------------------------------------
constexpr int kSortThreads = 256;
constexpr int kSortRounds = 16;
constexpr int kSortChunk = kSortThreads * kSortRounds;  // keys per block

// per-block digit counts, digit-major: hist[digit * nblocks + block]
__global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(const uint32_t *__restrict__ keys, uint64_t n, int shift, uint64_t *__restrict__ hist,
                  uint32_t nblocks)
{
    __shared__ uint32_t cnt[256];
    cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * kSortChunk;
#pragma unroll 4
    for (int r = 0; r < kSortRounds; ++r) {
        const uint64_t i = base + (uint64_t)r * kSortThreads + threadIdx.x;
        if (i < n) atomicAdd(&cnt[(keys[i] >> shift) & 0xffu], 1u);
    }
    __syncthreads();
    hist[(uint64_t)threadIdx.x * nblocks + blockIdx.x] = cnt[threadIdx.x];
}

// one CTA: in-place exclusive scan of m 64-bit counts
__global__ void __launch_bounds__(1024) scan_u64_kernel(uint64_t *a, uint64_t m)
{
    __shared__ uint64_t wsum[32];
    __shared__ uint64_t chunk_total;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint64_t carry = 0;
    for (uint64_t base = 0; base < m; base += 1024) {
        const uint64_t k = base + threadIdx.x;
        const uint64_t x = k < m ? a[k] : 0;
        uint64_t inc = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint64_t y = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += y;
        }
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        if (w == 0) {
            const uint64_t ws = wsum[lane];
            uint64_t winc = ws;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint64_t y = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += y;
            }
            wsum[lane] = winc - ws;
            if (lane == 31) chunk_total = winc;
        }
        __syncthreads();
        if (k < m) a[k] = carry + wsum[w] + inc - x;
        carry += chunk_total;
        __syncthreads();
    }
}

// Stable scatter: the block walks its chunk in rounds of 256 consecutive keys.  Inside a round a
// key's rank among equal digits is (matching lanes below it in its warp) + (counts of the lower
// warps); `running` carries the counts of the earlier rounds.
__global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(const uint32_t *__restrict__ keys_in, const int32_t *__restrict__ ids_in, uint64_t n,
                     int shift, const uint64_t *__restrict__ hist, uint32_t nblocks,
                     uint32_t *__restrict__ keys_out, int32_t *__restrict__ ids_out)
{
    __shared__ uint32_t running[256];
    __shared__ uint32_t cnt[kSortThreads / 32][256];
    __shared__ uint64_t gbase[256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    running[threadIdx.x] = 0;
    gbase[threadIdx.x] = hist[(uint64_t)threadIdx.x * nblocks + blockIdx.x];
#pragma unroll
    for (int q = 0; q < kSortThreads / 32; ++q) cnt[q][threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * kSortChunk;
    for (int r = 0; r < kSortRounds; ++r) {
        const uint64_t i = base + (uint64_t)r * kSortThreads + threadIdx.x;
        const bool live = i < n;
        uint32_t key = 0;
        int32_t id = 0;
        if (live) {
            key = keys_in[i];
            id = ids_in[i];
        }
        const uint32_t dg = live ? ((key >> shift) & 0xffu) : 0x100u;  // dead lanes match only each other
        const unsigned peers = __match_any_sync(0xffffffffu, dg);
        const uint32_t below = __popc(peers & ((1u << lane) - 1u));
        if (live && below == 0) cnt[w][dg] = __popc(peers);
        __syncthreads();
        if (live) {
            uint32_t off = running[dg] + below;
            for (int q = 0; q < w; ++q) off += cnt[q][dg];
            const uint64_t o = gbase[dg] + off;
            keys_out[o] = key;
            ids_out[o] = id;
        }
        __syncthreads();
        {
            uint32_t s = 0;
#pragma unroll
            for (int q = 0; q < kSortThreads / 32; ++q) {
                s += cnt[q][threadIdx.x];
                cnt[q][threadIdx.x] = 0;
            }
            running[threadIdx.x] += s;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
invert_perm_kernel(const int32_t *__restrict__ order, uint64_t n, int32_t *__restrict__ newid)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x)
        newid[order[i]] = (int32_t)i;
}

// first three slots of every simplex (two for dp1 == 2, the second repeated), renumbered; a vertex id
// outside [0, n) is passed through unchanged so that the CSR build still reports it
__global__ void __launch_bounds__(256)
relabel3_kernel(const int32_t *__restrict__ simplices, uint64_t S, int dp1, uint64_t n,
                const int32_t *__restrict__ newid, int32_t *__restrict__ out)
{
    for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < S;
         s += (uint64_t)gridDim.x * blockDim.x) {
        const int32_t *p = simplices + s * (uint64_t)dp1;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int32_t v = p[k < dp1 ? k : dp1 - 1];
            out[s * 3 + k] = ((uint64_t)(uint32_t)v < n && newid) ? newid[v] : v;
        }
    }
}

template <class T>
__global__ void __launch_bounds__(256)
gather_kernel(const T *__restrict__ src, const int32_t *__restrict__ idx, uint64_t n, T *__restrict__ dst)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x)
        dst[i] = src[idx[i]];
}
template <class T>
__global__ void __launch_bounds__(256)
scatter_kernel(const T *__restrict__ src, const int32_t *__restrict__ idx, uint64_t n, T *__restrict__ dst)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x)
        dst[idx[i]] = src[i];
}

static inline unsigned grid1d(uint64_t items, int per_sm)
{
    uint64_t b = (items + 255) / 256;
    const uint64_t cap = (uint64_t)sm_count() * per_sm;
    if (b > cap) b = cap;
    return (unsigned)(b < 1 ? 1 : b);
}

size_t radix_hist_bytes(uint64_t n)
{
    const uint64_t nblocks = (n + kSortChunk - 1) / kSortChunk;
    return align256r(256 * (nblocks + 1) * 8);
}

int radix_sort_pairs(uint32_t *k0, int32_t *i0, uint32_t *k1, int32_t *i1, uint64_t *hist, uint64_t n,
                     cudaStream_t st)
{
    if (n == 0) return SHGO_B200_OK;
    const uint32_t nblocks = (uint32_t)((n + kSortChunk - 1) / kSortChunk);
    // 4 passes: (k0, i0) -> (k1, i1) -> (k0, i0) -> (k1, i1) -> (k0, i0)
    uint32_t *kin = k0, *kout = k1;
    int32_t *iin = i0, *iout = i1;
    for (int pass = 0; pass < 4; ++pass) {
        radix_hist_kernel<<<nblocks, kSortThreads, 0, st>>>(kin, n, pass * 8, hist, nblocks);
        scan_u64_kernel<<<1, 1024, 0, st>>>(hist, 256ull * nblocks);
        radix_scatter_kernel<<<nblocks, kSortThreads, 0, st>>>(kin, iin, n, pass * 8, hist, nblocks, kout, iout);
        uint32_t *tk = kin;
        kin = kout, kout = tk;
        int32_t *ti = iin;
        iin = iout, iout = ti;
    }
    SHGO_CUDA_TRY(cudaGetLastError());
    return SHGO_B200_OK;
}

}  // namespace shgo_b200

using namespace shgo_b200;

extern "C" {

size_t shgo_b200_morton_workspace(uint64_t n)
{
    // keys x2, ids x1 (the other id buffer is `order`), histogram
    return 2 * align256r(n * 4) + align256r(n * 4) + radix_hist_bytes(n);
}

int shgo_b200_morton_order(const double *x_soa, uint64_t ld, int d, uint64_t n, const double *lb,
                           const double *ub, int32_t *order, int32_t *newid, void *tmp, size_t tmp_bytes,
                           void *stream)
{
    SHGO_REQUIRE(d >= 1, "d must be positive (got %d)", d);
    SHGO_REQUIRE(n <= 0x7fffffffull, "vertex count out of int32 range");
    if (n == 0) return SHGO_B200_OK;
    SHGO_REQUIRE(x_soa && lb && ub && order && tmp, "null pointer");
    SHGO_REQUIRE(ld >= n, "ld < n");
    if (tmp_bytes < shgo_b200_morton_workspace(n))
        return fail(SHGO_B200_ERR_WORKSPACE, "morton workspace: need %zu bytes, got %zu",
                    shgo_b200_morton_workspace(n), tmp_bytes);
    cudaStream_t st = as_stream(stream);
    unsigned char *p = static_cast<unsigned char *>(tmp);
    uint32_t *k0 = reinterpret_cast<uint32_t *>(p);
    p += align256r(n * 4);
    uint32_t *k1 = reinterpret_cast<uint32_t *>(p);
    p += align256r(n * 4);
    int32_t *i1 = reinterpret_cast<int32_t *>(p);
    p += align256r(n * 4);
    uint64_t *hist = reinterpret_cast<uint64_t *>(p);
    morton_key_kernel<<<grid1d(n, 8), 256, 0, st>>>(x_soa, ld, d, n, lb, ub, k0, order);
    if (int rc = radix_sort_pairs(k0, order, k1, i1, hist, n, st)) return rc;
    if (newid) invert_perm_kernel<<<grid1d(n, 8), 256, 0, st>>>(order, n, newid);
    SHGO_CUDA_TRY(cudaGetLastError());
    return SHGO_B200_OK;
}

int shgo_b200_relabel3(const int32_t *simplices, uint64_t S, int dp1, uint64_t n, const int32_t *newid,
                       int32_t *out, void *stream)
{
    SHGO_REQUIRE(dp1 >= 2, "simplices need at least two vertex slots (dp1 = %d)", dp1);
    if (S == 0) return SHGO_B200_OK;
    SHGO_REQUIRE(simplices && out, "null pointer");
    relabel3_kernel<<<grid1d(S, 8), 256, 0, as_stream(stream)>>>(simplices, S, dp1, n, newid, out);
    SHGO_CUDA_TRY(cudaGetLastError());
    return SHGO_B200_OK;
}

int shgo_b200_gather_f64(const double *src, const int32_t *idx, uint64_t n, double *dst, void *stream)
{
    if (n == 0) return SHGO_B200_OK;
    SHGO_REQUIRE(src && idx && dst, "null pointer");
    gather_kernel<double><<<grid1d(n, 8), 256, 0, as_stream(stream)>>>(src, idx, n, dst);
    SHGO_CUDA_TRY(cudaGetLastError());
    return SHGO_B200_OK;
}

int shgo_b200_gather_u8(const uint8_t *src, const int32_t *idx, uint64_t n, uint8_t *dst, void *stream)
{
    if (n == 0) return SHGO_B200_OK;
    SHGO_REQUIRE(src && idx && dst, "null pointer");
    gather_kernel<uint8_t><<<grid1d(n, 8), 256, 0, as_stream(stream)>>>(src, idx, n, dst);
    SHGO_CUDA_TRY(cudaGetLastError());
    return SHGO_B200_OK;
}

int shgo_b200_scatter_u8(const uint8_t *src, const int32_t *idx, uint64_t n, uint8_t *dst, void *stream)
{
    if (n == 0) return SHGO_B200_OK;
    SHGO_REQUIRE(src && idx && dst, "null pointer");
    scatter_kernel<uint8_t><<<grid1d(n, 8), 256, 0, as_stream(stream)>>>(src, idx, n, dst);
    SHGO_CUDA_TRY(cudaGetLastError());
    return SHGO_B200_OK;
}

}  // extern "C"
