// bounds.cu -- local bounds of the pool vertices.
//
// Reference: SHGO.construct_lcb_simplicial, shgo/_shgo.py:1246-1270.  For a minimiser v the local
// minimisation is boxed, per axis, by the coordinate of the neighbour nearest below / above v's own
// coordinate (the global bound where v has no neighbour on that side).  The reference walks v.nn
// as Python objects -- up to 59 048 of them for one vertex of the 10-D complex; here the pool's
// boxes come back as two (m, d) arrays and no neighbour is ever materialised on the host.
#include "common.cuh"

namespace shgo_b200 {

// explicit graph: one warp per pool vertex, lanes stride the neighbour list, one axis at a time
// (the list of a pool vertex is short and stays in L1 across the d passes)
__global__ void __launch_bounds__(256)
local_bounds_csr_kernel(const int64_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                        const double *__restrict__ x, uint64_t ld, int d,
                        const int64_t *__restrict__ pool, uint64_t m, const double *__restrict__ lb,
                        const double *__restrict__ ub, double *__restrict__ lo, double *__restrict__ hi)
{
    const int lane = threadIdx.x & 31;
    const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t q = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < m; q += nw) {
        const int64_t v = pool[q];
        const int64_t b = row_ptr[v], e = row_ptr[v + 1];
        for (int i = 0; i < d; ++i) {
            const double *xi = x + (uint64_t)i * ld;
            const double xv = xi[v];
            double l = lb[i], h = ub[i];
            for (int64_t k = b + lane; k < e; k += 32) {
                const double w = ld_nc_f64(xi + col[k]);
                if (w < xv && w > l) l = w;
                if (w > xv && w < h) h = w;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                l = fmax(l, __shfl_xor_sync(0xffffffffu, l, o));
                h = fmin(h, __shfl_xor_sync(0xffffffffu, h, o));
            }
            if (lane == 0) {
                lo[q * (uint64_t)d + i] = l;
                hi[q * (uint64_t)d + i] = h;
            }
        }
    }
}

// whole-generation hypercube complex: the neighbours of a vertex with lattice index p_i on axis i
// sit at p_i-2 .. p_i+2, and a neighbour at p_i-1 (p_i+1) exists whenever p_i > 0 (p_i < L): for a
// grid vertex the centroid of any adjacent cell on that side, for a centroid the corners of its own
// cell.  One thread per (pool vertex, axis).
__global__ void __launch_bounds__(256)
lattice_local_bounds_kernel(int d, int gen, const double *__restrict__ tab, uint64_t ngrid,
                            const int64_t *__restrict__ pool, uint64_t m, const double *__restrict__ lb,
                            const double *__restrict__ ub, double *__restrict__ lo, double *__restrict__ hi)
{
    const uint64_t K = 1ull << gen, L = 2 * K, L1 = L + 1;
    const uint64_t total = m * (uint64_t)d;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t q = t / (uint64_t)d;
        const int i = (int)(t - q * (uint64_t)d);
        uint64_t id = (uint64_t)pool[q], p;
        if (id < ngrid) {
            for (int j = 0; j < i; ++j) id /= K + 1;
            p = 2 * (id % (K + 1));
        } else {
            id -= ngrid;
            p = 2 * ((id >> ((uint64_t)gen * i)) & (K - 1)) + 1;
        }
        const double *row = tab + (uint64_t)i * L1;
        double l = lb[i], h = ub[i];
        if (p > 0 && row[p - 1] > l) l = row[p - 1];
        if (p < L && row[p + 1] < h) h = row[p + 1];
        lo[t] = l;
        hi[t] = h;
    }
}

}  // namespace shgo_b200

using namespace shgo_b200;

extern "C" int shgo_b200_local_bounds_csr(const int64_t *row_ptr, const int32_t *col, const double *x_soa,
                                          uint64_t ld, int d, const int64_t *pool, uint64_t m,
                                          const double *lb, const double *ub, double *lo, double *hi,
                                          void *stream)
{
    SHGO_REQUIRE(d >= 1, "d must be positive (got %d)", d);
    if (m == 0) return SHGO_B200_OK;
    // col may be null: a graph without edges (a single isolated vertex) has no column array
    SHGO_REQUIRE(row_ptr && x_soa && pool && lb && ub && lo && hi, "null pointer");
    uint64_t blocks = (m + 7) / 8;  // 8 warps per block
    const uint64_t cap = (uint64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    local_bounds_csr_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(row_ptr, col, x_soa, ld, d, pool, m,
                                                                            lb, ub, lo, hi);
    SHGO_CUDA_TRY(cudaGetLastError());
    return SHGO_B200_OK;
}

extern "C" int shgo_b200_lattice_local_bounds(int d, int gen, const double *tab, const int64_t *pool,
                                              uint64_t m, const double *lb, const double *ub, double *lo,
                                              double *hi, void *stream)
{
    SHGO_REQUIRE(d >= 1 && d <= 12, "lattice supports 1 <= d <= 12 (got %d)", d);
    SHGO_REQUIRE(gen >= 0 && gen <= 20, "generation %d outside 0..20", gen);
    double ng = 1, nc = 1;
    uint64_t ngrid = 1;
    for (int i = 0; i < d; ++i) {
        ng *= (double)((1u << gen) + 1);
        nc *= (double)(1u << gen);
        ngrid *= (1ull << gen) + 1;
    }
    SHGO_REQUIRE(ng + nc < 4294967296.0, "lattice with more than 2^32 vertices is not supported");
    if (m == 0) return SHGO_B200_OK;
    SHGO_REQUIRE(tab && pool && lb && ub && lo && hi, "null pointer");
    uint64_t blocks = (m * (uint64_t)d + 255) / 256;
    const uint64_t cap = (uint64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    lattice_local_bounds_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(d, gen, tab, ngrid, pool, m, lb,
                                                                                ub, lo, hi);
    SHGO_CUDA_TRY(cudaGetLastError());
    return SHGO_B200_OK;
}
