// lattice.cu -- K5: star test over the whole-generation hypercube complex with IMPLICIT adjacency.
//
// The reference grows this complex edge by edge in Python (Complex.triangulate / refine_all,
// _complex.py:192-472, 514-976) and then runs VertexScalarField.minimiser over explicit neighbour
// sets (_vertex.py:144-151).  A 10-D grid vertex has up to 3^10-1 = 59 048 neighbours, so an
// explicit CSR is out of the question (~1e10 entries at generation 2); the closed form of
// SURVEY.md Appendix B lets a warp enumerate them on the fly:
//   centroid 2c+1        <-> the 2^d corners of its cell
//   grid vertex k        <-> centroids of the <= 2^d cells around it
//   grid vertex k        <-> k + sum_{i in S} s_i e_i for every non-empty S inside the odd-index
//                            axes or inside the even-index axes, s_i = +-1, |S| < d, in range
// One WARP owns one vertex: lane l tests neighbour numbers l, l+32, ... and the warp votes after
// every batch of 32, leaving as soon as some neighbour is not strictly larger (the reference's
// all() short-circuits the same way).  Vertex ids: see include/shgo_b200.h.
#include "common.cuh"

namespace shgo_b200 {

template <int D>
__global__ void __launch_bounds__(256)
lattice_star_kernel(int gen, const double *__restrict__ f, uint64_t v0, uint64_t nv, uint64_t ngrid,
                    uint8_t *__restrict__ is_min)
{
    const int K = 1 << gen, K1 = K + 1;
    const int lane = threadIdx.x & 31;
    const uint64_t gwarp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    int64_t gstride[D], cstride[D];
    {
        int64_t g = 1, c = 1;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            gstride[i] = g;
            cstride[i] = c;
            g *= K1;
            c *= K;
        }
    }
    constexpr unsigned kFull = 0xffffffffu;
    constexpr int kCorners = 1 << D;
    for (uint64_t lv = gwarp; lv < nv; lv += nwarps) {
        const uint64_t v = v0 + lv;
        const double fv = ld_nc_f64(f + v);
        bool fail = false;
        int k[D];
        if (v >= ngrid) {
            uint64_t r = v - ngrid;
            int64_t base = 0;
#pragma unroll
            for (int i = 0; i < D; ++i) {
                k[i] = (int)(r & (uint64_t)(K - 1));
                r >>= gen;
                base += k[i] * gstride[i];
            }
            for (int m0 = 0; m0 < kCorners && !fail; m0 += 32) {
                const int m = m0 + lane;
                bool bad = false;
                if (m < kCorners) {
                    int64_t nb = base;
#pragma unroll
                    for (int i = 0; i < D; ++i)
                        if ((m >> i) & 1) nb += gstride[i];
                    bad = !(fv < ld_nc_f64(f + nb));
                }
                fail = __any_sync(kFull, bad);
            }
        } else {
            uint64_t r = v;
            unsigned oddmask = 0;
#pragma unroll
            for (int i = 0; i < D; ++i) {
                k[i] = (int)(r % (uint64_t)K1);
                r /= (uint64_t)K1;
                oddmask |= (unsigned)(k[i] & 1) << i;
            }
            // centroids of the cells around the vertex
            for (int m0 = 0; m0 < kCorners && !fail; m0 += 32) {
                const int m = m0 + lane;
                bool bad = false;
                if (m < kCorners) {
                    int64_t nb = (int64_t)ngrid;
                    bool ok = true;
#pragma unroll
                    for (int i = 0; i < D; ++i) {
                        int c = k[i] - ((m >> i) & 1);
                        ok = ok && c >= 0 && c < K;
                        nb += c * cstride[i];
                    }
                    if (ok) bad = !(fv < ld_nc_f64(f + nb));
                }
                fail = __any_sync(kFull, bad);
            }
            // grid-grid moves inside the odd-index axes, then inside the even-index axes
#pragma unroll 1
            for (int cls = 0; cls < 2 && !fail; ++cls) {
                const unsigned amask = (cls == 0 ? oddmask : ~oddmask) & ((1u << D) - 1u);
                const int na = __popc(amask);
                if (na == 0) continue;
                unsigned combos = 1;
                for (int i = 0; i < na; ++i) combos *= 3u;
                for (unsigned t0 = 1; t0 < combos && !fail; t0 += 32) {
                    unsigned t = t0 + lane;
                    bool bad = false;
                    if (t < combos) {
                        int64_t nb = (int64_t)v;
                        int moved = 0;
                        bool ok = true;
#pragma unroll
                        for (int i = 0; i < D; ++i) {
                            if ((amask >> i) & 1u) {
                                unsigned q = t / 3u, dg = t - q * 3u;
                                t = q;
                                if (dg) {
                                    ++moved;
                                    int kk = k[i] + (dg == 1 ? 1 : -1);
                                    ok = ok && kk >= 0 && kk <= K;
                                    nb += dg == 1 ? gstride[i] : -gstride[i];
                                }
                            }
                        }
                        if (ok && moved < D) bad = !(fv < ld_nc_f64(f + nb));
                    }
                    fail = __any_sync(kFull, bad);
                }
            }
        }
        if (lane == 0) is_min[lv] = fail ? 0 : 1;
    }
}

template <int D>
static int launch_lattice_star(int gen, const double *f, uint64_t v0, uint64_t nv, uint64_t ngrid,
                               uint8_t *is_min, cudaStream_t st)
{
    uint64_t blocks = (nv + 7) / 8;  // 8 warps per CTA, one vertex per warp per pass
    uint64_t cap = (uint64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    lattice_star_kernel<D><<<(unsigned)blocks, 256, 0, st>>>(gen, f, v0, nv, ngrid, is_min);
    SHGO_CUDA_TRY(cudaGetLastError());
    return SHGO_B200_OK;
}

}  // namespace shgo_b200

using namespace shgo_b200;

extern "C" int shgo_b200_lattice_star(int d, int gen, const double *f, uint64_t v0, uint64_t nv,
                                      uint8_t *is_min, void *stream)
{
    SHGO_REQUIRE(f && is_min, "null pointer");
    SHGO_REQUIRE(d >= 1 && d <= 12, "lattice star test supports 1 <= d <= 12 (got %d)", d);
    SHGO_REQUIRE(gen >= 0 && gen <= 20, "generation %d outside 0..20", gen);
    double ng = 1, nc = 1;
    uint64_t ngrid = 1, ncent = 1;
    for (int i = 0; i < d; ++i) {
        ng *= (double)((1u << gen) + 1);
        nc *= (double)(1u << gen);
        ngrid *= (1ull << gen) + 1;
        ncent *= 1ull << gen;
    }
    SHGO_REQUIRE(ng + nc < 4294967296.0, "lattice with more than 2^32 vertices is not supported");
    SHGO_REQUIRE(v0 + nv <= ngrid + ncent, "vertex range exceeds the lattice");
    if (nv == 0) return SHGO_B200_OK;
    cudaStream_t st = as_stream(stream);
    switch (d) {
#define SHGO_D(DD) \
    case DD: return launch_lattice_star<DD>(gen, f, v0, nv, ngrid, is_min, st);
        SHGO_D(1) SHGO_D(2) SHGO_D(3) SHGO_D(4) SHGO_D(5) SHGO_D(6)
        SHGO_D(7) SHGO_D(8) SHGO_D(9) SHGO_D(10) SHGO_D(11) SHGO_D(12)
#undef SHGO_D
    }
    return fail(SHGO_B200_ERR_ARG, "unreachable");
}
