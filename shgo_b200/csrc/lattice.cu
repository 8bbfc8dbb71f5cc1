// lattice.cu -- K5: star test over the whole-generation hypercube complex with IMPLICIT adjacency.
//
// The reference grows this complex edge by edge in Python (Complex.triangulate / refine_all,
// _complex.py:192-472, 514-976) and then runs VertexScalarField.minimiser over explicit neighbour
// sets (_vertex.py:144-151).  A 10-D grid vertex has up to 3^10-1 = 59 048 neighbours, so an
// explicit CSR is out of the question (~1e10 entries at generation 2); the closed form of
// SURVEY.md Appendix B lets the threads enumerate them on the fly:
//   centroid 2c+1        <-> the 2^d corners of its cell
//   grid vertex k        <-> centroids of the <= 2^d cells around it
//   grid vertex k        <-> k + sum_{i in S} s_i e_i for every non-empty S inside the odd-index
//                            axes or inside the even-index axes, s_i = +-1, |S| < d, in range
// The work per vertex is wildly uneven -- almost every vertex is rejected by its first few
// neighbours, a true minimiser needs all 59 048 -- so the test runs as a funnel:
//   A. one THREAD per vertex: the <= 2d single-axis moves (coalesced gathers) / 2d corners;
//   B. survivors (warp ballot): the whole WARP tests the next kWarpBatches*32 neighbours;
//   C. what is still undecided (a fraction of a percent) is compacted and finished by a whole CTA
//      per vertex, 4 neighbours per thread and vote, so that a 59 048-neighbour star costs ~58
//      latency round trips instead of 1845 and the heavy vertices spread over all SMs.
// Every stage leaves as soon as some neighbour is not strictly larger (the reference's all()
// short-circuits the same way).  Vertex ids: see include/shgo_b200.h.
#include "common.cuh"

namespace shgo_b200 {

constexpr int kWarpBatches = 4;   // stage B: 128 neighbours
constexpr int kHeavyThreads = 256;
constexpr int kHeavyIlp = 4;

template <int D>
struct LatticeGeom {
    uint32_t gstride[D], cstride[D];
    uint32_t ngrid;
    int gen, K, K1;
    __device__ LatticeGeom(int gen_, uint32_t ngrid_) : ngrid(ngrid_), gen(gen_), K(1 << gen_), K1((1 << gen_) + 1)
    {
        uint32_t g = 1, c = 1;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            gstride[i] = g;
            cstride[i] = c;
            g *= (uint32_t)K1;
            c *= (uint32_t)K;
        }
    }
};

// The star of vertex v as a numbered sequence 0 .. total-1 that a GROUP of threads walks together.
template <int D>
struct StarWalk {
    const LatticeGeom<D> &g;
    int k[D];
    uint32_t v, base;        // base: id of corner 0 (centroid) / unused (grid vertex)
    unsigned amask[2];       // odd-index axes, even-index axes (grid vertex)
    unsigned combos[2];      // 3^|axes| per class
    bool centroid;

    __device__ StarWalk(const LatticeGeom<D> &geom, uint32_t v_) : g(geom), v(v_)
    {
        centroid = v >= g.ngrid;
        if (centroid) {
            uint32_t r = v - g.ngrid;
            base = 0;
#pragma unroll
            for (int i = 0; i < D; ++i) {
                k[i] = (int)(r & (uint32_t)(g.K - 1));
                r >>= g.gen;
                base += (uint32_t)k[i] * g.gstride[i];
            }
            amask[0] = amask[1] = 0;
            combos[0] = combos[1] = 1;
        } else {
            uint32_t r = v;
            unsigned odd = 0;
#pragma unroll
            for (int i = 0; i < D; ++i) {
                const uint32_t q = r / (uint32_t)g.K1;
                k[i] = (int)(r - q * (uint32_t)g.K1);
                r = q;
                odd |= (unsigned)(k[i] & 1) << i;
            }
            amask[0] = odd;
            amask[1] = ~odd & ((1u << D) - 1u);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                unsigned n = 1;
                for (int i = __popc(amask[c]); i > 0; --i) n *= 3u;
                combos[c] = n;
            }
        }
    }
    // segments: [0, 2^D) corners / surrounding cells; then class-0 moves 1..combos0-1; class-1 moves
    __device__ __forceinline__ unsigned total() const
    {
        return (1u << D) + (centroid ? 0u : (combos[0] - 1) + (combos[1] - 1));
    }
    // id of neighbour number t, or false when that slot is not a neighbour (out of range, |S| = d)
    __device__ __forceinline__ bool neighbour(unsigned t, uint32_t &nb) const
    {
        if (t < (1u << D)) {
            if (centroid) {
                nb = base;
#pragma unroll
                for (int i = 0; i < D; ++i)
                    if ((t >> i) & 1u) nb += g.gstride[i];
                return true;
            }
            nb = g.ngrid;
            bool ok = true;
#pragma unroll
            for (int i = 0; i < D; ++i) {
                const int c = k[i] - (int)((t >> i) & 1u);
                ok = ok && c >= 0 && c < g.K;
                nb += (uint32_t)c * g.cstride[i];
            }
            return ok;
        }
        t -= (1u << D);
        int cls = 0;
        if (t >= combos[0] - 1) {
            t -= combos[0] - 1;
            cls = 1;
        }
        t += 1;  // move numbers start at 1 (0 = stay)
        const unsigned am = amask[cls];
        nb = v;
        int moved = 0;
        bool ok = true;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            if ((am >> i) & 1u) {
                const unsigned q = t / 3u, dg = t - q * 3u;
                t = q;
                if (dg) {
                    ++moved;
                    const int kk = k[i] + (dg == 1 ? 1 : -1);
                    ok = ok && kk >= 0 && kk <= g.K;
                    nb = dg == 1 ? nb + g.gstride[i] : nb - g.gstride[i];
                }
            }
        }
        return ok && moved < D;
    }
};

// flags: 0 = not a minimiser, 1 = minimiser, 2 = undecided after stages A+B
template <int D>
__global__ void __launch_bounds__(256)
lattice_star_kernel(int gen, const double *__restrict__ f, uint64_t v0, uint64_t nv, uint64_t ngrid64,
                    uint8_t *__restrict__ flags, uint32_t part, uint32_t nparts, int granule_log2)
{
    const LatticeGeom<D> g(gen, (uint32_t)ngrid64);
    const int lane = threadIdx.x & 31;
    const uint64_t gwarp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    constexpr unsigned kFull = 0xffffffffu;
    constexpr int kQuick = 2 * D < (1 << D) ? 2 * D : (1 << D);
    for (uint64_t base = gwarp << 5; base < nv; base += nwarps << 5) {
        // multi-GPU: only the granules dealt to this part (the others' flags were zeroed)
        if (nparts > 1 && ((v0 + base) >> granule_log2) % nparts != part) continue;
        const uint64_t lv = base + lane;
        const bool valid = lv < nv;
        const uint32_t v = (uint32_t)(v0 + (valid ? lv : 0));
        const double fv = ld_nc_f64(f + v);
        // ---- stage A ----
        bool alive = valid;
        if (v >= g.ngrid) {
            uint32_t r = v - g.ngrid, cb = 0;
#pragma unroll
            for (int i = 0; i < D; ++i) {
                cb += (r & (uint32_t)(g.K - 1)) * g.gstride[i];
                r >>= gen;
            }
#pragma unroll
            for (int m = 0; m < kQuick; ++m) {
                uint32_t nb = cb;
#pragma unroll
                for (int i = 0; i < D; ++i)
                    if ((m >> i) & 1) nb += g.gstride[i];
                if (alive) alive = fv < ld_nc_f64(f + nb);
            }
        } else if (D > 1) {
            uint32_t r = v;
#pragma unroll
            for (int i = 0; i < D; ++i) {
                const uint32_t q = r / (uint32_t)g.K1;
                const int ki = (int)(r - q * (uint32_t)g.K1);
                r = q;
                if (alive && ki < g.K) alive = fv < ld_nc_f64(f + v + g.gstride[i]);
                if (alive && ki > 0) alive = fv < ld_nc_f64(f + v - g.gstride[i]);
            }
        }
        // ---- stage B ----
        uint8_t flag = alive ? 1 : 0;
        unsigned todo = __ballot_sync(kFull, alive);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const StarWalk<D> w(g, __shfl_sync(kFull, v, src));
            const double fvv = __shfl_sync(kFull, fv, src);
            const unsigned total = w.total();
            bool fail = false;
            unsigned t0 = 0;
            for (int b = 0; b < kWarpBatches && t0 < total && !fail; ++b, t0 += 32) {
                uint32_t nb;
                const unsigned t = t0 + lane;
                const bool bad = t < total && w.neighbour(t, nb) && !(fvv < ld_nc_f64(f + nb));
                fail = __any_sync(kFull, bad);
            }
            if (lane == src) flag = fail ? 0 : (t0 < total ? 2 : 1);
        }
        if (valid) flags[lv] = flag;
    }
}

// ---- stage C: one CTA per undecided vertex ----
// Corner / cell slots are decoded directly (at most 2^d of them).  The grid-grid moves -- up to
// 3^d - 1 per class -- are walked with a base-3 ODOMETER: every thread owns a contiguous range of
// move numbers, decodes its first one once and then steps (+1 with carry), keeping the neighbour
// id, the number of moved axes and the number of out-of-range axes incrementally: ~12 instructions
// per neighbour instead of ~150 for a full decode.  Per-position tables (stride, at-lower-bound,
// at-upper-bound of the p-th axis of the class) live in shared memory.
template <int D>
__global__ void __launch_bounds__(kHeavyThreads)
lattice_star_heavy_kernel(int gen, const double *__restrict__ f, uint64_t v0, uint64_t ngrid64,
                          const int64_t *__restrict__ cand, const uint64_t *__restrict__ ncand,
                          uint8_t *__restrict__ flags)
{
    __shared__ uint32_t s_gs[D];
    __shared__ int s_lo[D], s_hi[D];
    const LatticeGeom<D> g(gen, (uint32_t)ngrid64);
    const uint64_t n = *ncand;
    const int tid = threadIdx.x;
    for (uint64_t c = blockIdx.x; c < n; c += gridDim.x) {
        const uint64_t lv = (uint64_t)cand[c];
        const uint32_t v = (uint32_t)(v0 + lv);
        const double fv = ld_nc_f64(f + v);
        const StarWalk<D> w(g, v);
        int fail = 0;
        // corners (centroid) / surrounding cells (grid vertex): slots kWarpBatches*32 .. 2^D - 1
        for (unsigned t0 = kWarpBatches * 32; t0 < (1u << D) && !fail; t0 += kHeavyThreads * kHeavyIlp) {
            uint32_t nb[kHeavyIlp];
            bool has[kHeavyIlp];
            double fn[kHeavyIlp];
#pragma unroll
            for (int u = 0; u < kHeavyIlp; ++u) {
                const unsigned t = t0 + u * kHeavyThreads + tid;
                has[u] = t < (1u << D) && w.neighbour(t, nb[u]);
            }
#pragma unroll
            for (int u = 0; u < kHeavyIlp; ++u) fn[u] = has[u] ? ld_nc_f64(f + nb[u]) : pos_inf();
            bool bad = false;
#pragma unroll
            for (int u = 0; u < kHeavyIlp; ++u) bad = bad || !(fv < fn[u]);
            fail = __syncthreads_or(bad);
        }
        // grid-grid moves (stage B only reached into them when 2^D < kWarpBatches*32; re-testing a
        // few slots is harmless)
        for (int cls = 0; cls < 2 && !fail && !w.centroid; ++cls) {
            const unsigned am = w.amask[cls];
            const int na = __popc(am);
            if (na == 0) continue;
            __syncthreads();
            if (tid == 0) {
                int p = 0;
#pragma unroll
                for (int i = 0; i < D; ++i)
                    if ((am >> i) & 1u) {
                        s_gs[p] = g.gstride[i];
                        s_lo[p] = w.k[i] == 0;
                        s_hi[p] = w.k[i] == g.K;
                        ++p;
                    }
            }
            __syncthreads();
            const unsigned total = w.combos[cls] - 1;                       // move numbers 1 .. total
            const unsigned chunk = (total + kHeavyThreads - 1) / kHeavyThreads;
            const unsigned begin = 1 + (unsigned)tid * chunk;
            const unsigned end = min(begin + chunk, total + 1);
            // odometer state of move number `begin`
            uint32_t digits = 0, nb = v;
            int moved = 0, oob = 0;
            {
                unsigned tt = begin <= total ? begin : 0;
                for (int p = 0; p < na; ++p) {
                    const unsigned q = tt / 3u, dg = tt - q * 3u;
                    tt = q;
                    digits |= dg << (2 * p);
                    if (dg == 1) { nb += s_gs[p]; ++moved; oob += s_hi[p]; }
                    if (dg == 2) { nb -= s_gs[p]; ++moved; oob += s_lo[p]; }
                }
            }
            for (unsigned j = 0; j < chunk && !fail; j += kHeavyIlp) {
                uint32_t nbs[kHeavyIlp];
                bool has[kHeavyIlp];
                double fn[kHeavyIlp];
#pragma unroll
                for (int u = 0; u < kHeavyIlp; ++u) {
                    const unsigned cur = begin + j + u;
                    has[u] = cur < end && oob == 0 && moved < D;
                    nbs[u] = nb;
                    if (cur + 1 < end) {  // step the odometer: +1 with carry
                        int p = 0;
                        while (true) {
                            const uint32_t dg = (digits >> (2 * p)) & 3u;
                            if (dg == 0) { digits += 1u << (2 * p); nb += s_gs[p]; ++moved; oob += s_hi[p]; break; }
                            if (dg == 1) { digits += 1u << (2 * p); nb -= 2 * s_gs[p]; oob += s_lo[p] - s_hi[p]; break; }
                            digits &= ~(3u << (2 * p));
                            nb += s_gs[p];
                            --moved;
                            oob -= s_lo[p];
                            ++p;
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < kHeavyIlp; ++u) fn[u] = has[u] ? ld_nc_f64(f + nbs[u]) : pos_inf();
                bool bad = false;
#pragma unroll
                for (int u = 0; u < kHeavyIlp; ++u) bad = bad || !(fv < fn[u]);
                fail = __syncthreads_or(bad);
            }
        }
        if (tid == 0) flags[lv] = fail ? 0 : 1;
        __syncthreads();
    }
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

template <int D>
static int launch_lattice_star(int gen, const double *f, uint64_t v0, uint64_t nv, uint64_t ngrid,
                               uint8_t *flags, void *tmp, cudaStream_t st, int part, int nparts, int granule_log2)
{
    if (nparts > 1) SHGO_CUDA_TRY(cudaMemsetAsync(flags, 0, nv, st));
    uint64_t blocks = (nv + 255) / 256;  // 8 warps per CTA, 32 vertices per warp per pass
    uint64_t cap = (uint64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    lattice_star_kernel<D><<<(unsigned)blocks, 256, 0, st>>>(gen, f, v0, nv, ngrid, flags, (uint32_t)part,
                                                             (uint32_t)nparts, granule_log2);
    // stage C on the undecided vertices
    unsigned char *p = static_cast<unsigned char *>(tmp);
    int64_t *cand = reinterpret_cast<int64_t *>(p);
    p += align256(nv * sizeof(int64_t));
    void *ctmp = p;
    p += shgo_b200_compact_workspace(nv);
    uint64_t *ncand = reinterpret_cast<uint64_t *>(p);
    if (int rc = compact_flags(flags, nv, /*match=*/2, 0, cand, nv, ncand, ctmp, st)) return rc;
    lattice_star_heavy_kernel<D><<<sm_count() * 8, kHeavyThreads, 0, st>>>(gen, f, v0, ngrid, cand, ncand, flags);
    SHGO_CUDA_TRY(cudaGetLastError());
    return SHGO_B200_OK;
}

}  // namespace shgo_b200

using namespace shgo_b200;

extern "C" size_t shgo_b200_lattice_star_workspace(uint64_t nv)
{
    return align256(nv * sizeof(int64_t)) + shgo_b200_compact_workspace(nv) + align256(16);
}

extern "C" int shgo_b200_lattice_star(int d, int gen, const double *f, uint64_t v0, uint64_t nv,
                                      uint8_t *is_min, void *tmp, size_t tmp_bytes, void *stream)
{
    return shgo_b200_lattice_star_part(d, gen, f, v0, nv, 0, 1, SHGO_B200_LATTICE_MIN_GRANULE, is_min, tmp, tmp_bytes,
                                       stream);
}

extern "C" int shgo_b200_lattice_star_part(int d, int gen, const double *f, uint64_t v0, uint64_t nv,
                                           int part, int nparts, uint32_t granule, uint8_t *is_min,
                                           void *tmp, size_t tmp_bytes, void *stream)
{
    SHGO_REQUIRE(nparts >= 1 && part >= 0 && part < nparts, "bad part %d of %d", part, nparts);
    SHGO_REQUIRE(granule >= 32 && (granule & (granule - 1)) == 0, "granule must be a power of two >= 32");
    int glog = 0;
    while ((1u << glog) < granule) ++glog;
    SHGO_REQUIRE(nparts == 1 || v0 % 32 == 0, "v0 must be a multiple of 32 when the range is dealt out in parts");
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
    SHGO_REQUIRE(f && is_min && tmp, "null pointer");
    if (tmp_bytes < shgo_b200_lattice_star_workspace(nv))
        return fail(SHGO_B200_ERR_WORKSPACE, "lattice star workspace: need %zu bytes, got %zu",
                    shgo_b200_lattice_star_workspace(nv), tmp_bytes);
    cudaStream_t st = as_stream(stream);
    switch (d) {
#define SHGO_D(DD) \
    case DD: return launch_lattice_star<DD>(gen, f, v0, nv, ngrid, is_min, tmp, st, part, nparts, glog);
        SHGO_D(1) SHGO_D(2) SHGO_D(3) SHGO_D(4) SHGO_D(5) SHGO_D(6)
        SHGO_D(7) SHGO_D(8) SHGO_D(9) SHGO_D(10) SHGO_D(11) SHGO_D(12)
#undef SHGO_D
    }
    return fail(SHGO_B200_ERR_ARG, "unreachable");
}
