// sources.cuh -- where the coordinates of a point come from.
//
// Every evaluation kernel is written against a "point source": an object that hands a thread
// the j-th coordinate of its P points.  Three sources exist:
//   SobolSrc    points are never stored: coordinate j of Sobol index k is rebuilt from the
//               direction numbers (reference: qmc.Sobol + bounds scaling, _shgo.py:703-711,
//               1441-1449)
//   SoaSrc      caller-supplied points, coordinate-major [d][ld] (coalesced 8-byte loads)
//   LatticeSrc  vertices of the whole-generation hypercube complex, coordinate = table lookup
//               (reference: Complex.triangulate/refine_all, _complex.py:192-472,514-976)
// A thread owns points  tile_first + i*256 + threadIdx.x,  i = 0..P-1, so every per-i access
// of a warp is one contiguous 256-byte segment.
// load(j, x) must be called with j ascending from 0 within a tile (LatticeSrc peels digits).
#pragma once
#include "common.cuh"

namespace shgo_b200 {

// ---------------------------------------------------------------------------------------------
// Sobol: k = 256*M + r, r = threadIdx.x (the tile base is 256*P aligned).  With
// G = k ^ (k>>1) (Gray code) the 30-bit integer coordinate is the GF(2)-linear form
//   u_j(k) = LOW_j[r] ^ AH_j(M_tile) ^ AI_j[i]
//   LOW_j[r] = xor_{b<8, bit b of r^(r>>1)} sv_j[b]                (256 entries per axis)
//   w_j[c]   = sv_j[c+7] ^ sv_j[c+8]                               (how bit c of M enters G)
//   AH_j(M)  = xor_{c : M_c} w_j[c]   for the tile's M (multiple of P), once per tile per axis
//   AI_j[i]  = AH_j(i), i < P                                      (M_tile + i = M_tile | i)
// so one point-coordinate costs one broadcast LDS + one XOR, no per-thread generator state.
// The scaling replicates numpy bit for bit: x = fl(fl(u*2^-30 * (ub-lb)) + lb).  The int->double
// conversion and the exact 2^-30 are folded into one DFMA: with D = 2^52 + u (bit pattern
// 0x43300000:u) fma(D, rs, -2^52*rs) = fl(u*rs), rs = (ub-lb)*2^-30 (power-of-two scaling is
// exact), which equals numpy's fl((u*2^-30)*(ub-lb)).
// ---------------------------------------------------------------------------------------------
template <int P>
struct SobolSrc {
    struct Params {
        const uint32_t *sv;
        const double *lb;
        const double *ub;
        uint64_t k0;
        uint64_t n;
        int d;
    };
    static constexpr int kP = P;
    static constexpr int kTile = kEvalThreads * P;
    static_assert((P & (P - 1)) == 0, "P must be a power of two");

    uint32_t *low, *whi, *ai, *ah;
    double *rs, *cc, *lbs;
    uint64_t k0, n, kb, tile_first;
    int d;

    static size_t smem_bytes(int d)
    {
        size_t words = (size_t)d * (256 + 22 + P + 1);
        words = (words + 1) & ~(size_t)1;
        return words * 4 + (size_t)d * 3 * 8;
    }

    __device__ SobolSrc(const Params &p, unsigned char *smem) : k0(p.k0), n(p.n), d(p.d)
    {
        low = reinterpret_cast<uint32_t *>(smem);
        whi = low + d * 256;
        ai = whi + d * 22;
        ah = ai + d * P;
        size_t words = ((size_t)d * (256 + 22 + P + 1) + 1) & ~(size_t)1;
        rs = reinterpret_cast<double *>(smem + words * 4);
        cc = rs + d;
        lbs = cc + d;
        kb = k0 & ~(uint64_t)(kTile - 1);
        const int t = threadIdx.x;
        const uint32_t g = (uint32_t)t ^ ((uint32_t)t >> 1);
        for (int j = 0; j < d; ++j) {
            uint32_t v = 0;
#pragma unroll
            for (int b = 0; b < 8; ++b)
                if ((g >> b) & 1u) v ^= __ldg(p.sv + j * kSobolBits + b);
            low[j * 256 + t] = v;
        }
        for (int e = t; e < d * 22; e += kEvalThreads) {
            int j = e / 22, c = e - j * 22;
            uint32_t v = __ldg(p.sv + j * kSobolBits + c + 7);
            if (c + 8 < kSobolBits) v ^= __ldg(p.sv + j * kSobolBits + c + 8);
            whi[e] = v;
        }
        if (t < d) {
            double r = __dsub_rn(__ldg(p.ub + t), __ldg(p.lb + t));
            double s = r * 9.313225746154785e-10;  // 2^-30, exact scaling
            rs[t] = s;
            cc[t] = -(s * 4503599627370496.0);  // -2^52 * rs, exact
            lbs[t] = __ldg(p.lb + t);
        }
        __syncthreads();
        for (int e = t; e < d * P; e += kEvalThreads) {
            int j = e / P, i = e - j * P;
            uint32_t v = 0;
            for (int b = 0; (i >> b) != 0; ++b)
                if ((i >> b) & 1) v ^= whi[j * 22 + b];
            ai[e] = v;
        }
        __syncthreads();
    }

    __device__ uint64_t num_tiles() const { return (k0 + n - kb + kTile - 1) / kTile; }

    __device__ void begin_tile(uint64_t t)
    {
        __syncthreads();  // everyone is done with the previous tile's AH
        tile_first = kb + t * (uint64_t)kTile;
        if ((int)threadIdx.x < d) {
            const int j = threadIdx.x;
            uint32_t M = (uint32_t)(tile_first >> 8), a = 0;
            for (int c = 0; M != 0; ++c, M >>= 1)
                if (M & 1u) a ^= whi[j * 22 + c];
            ah[j] = a;
        }
        __syncthreads();
    }

    __device__ __forceinline__ void load(int j, double (&x)[P])
    {
        const uint32_t base = low[j * 256 + threadIdx.x] ^ ah[j];
        const double s = rs[j], c = cc[j], l = lbs[j];
#pragma unroll
        for (int i = 0; i < P; ++i) {
            uint32_t u = base ^ ai[j * P + i];
            double D = __hiloint2double(0x43300000, (int)u);
            x[i] = __dadd_rn(__fma_rn(D, s, c), l);
        }
    }

    __device__ __forceinline__ bool valid(int i, uint64_t &out) const
    {
        uint64_t k = tile_first + (uint64_t)i * kEvalThreads + threadIdx.x;
        out = k - k0;
        return k >= k0 && k < k0 + n;
    }
    // whole tile inside the requested range?  out = output index of the tile's first point
    __device__ __forceinline__ bool full_tile(uint64_t &out) const
    {
        out = tile_first - k0;
        return tile_first >= k0 && tile_first + kTile <= k0 + n;
    }
};

// ---------------------------------------------------------------------------------------------
template <int P>
struct SoaSrc {
    struct Params {
        const double *x;
        uint64_t ld;
        uint64_t n;
        int d;
    };
    static constexpr int kP = P;
    static constexpr int kTile = kEvalThreads * P;
    const double *xs;
    uint64_t ld, n, tile_first;
    int d;

    static size_t smem_bytes(int) { return 0; }
    __device__ SoaSrc(const Params &p, unsigned char *) : xs(p.x), ld(p.ld), n(p.n), d(p.d) {}
    __device__ uint64_t num_tiles() const { return (n + kTile - 1) / kTile; }
    __device__ void begin_tile(uint64_t t) { tile_first = t * (uint64_t)kTile; }
    __device__ __forceinline__ void load(int j, double (&x)[P])
    {
#pragma unroll
        for (int i = 0; i < P; ++i) {
            uint64_t k = tile_first + (uint64_t)i * kEvalThreads + threadIdx.x;
            x[i] = k < n ? ld_nc_f64(xs + (uint64_t)j * ld + k) : 0.0;
        }
    }
    __device__ __forceinline__ bool valid(int i, uint64_t &out) const
    {
        out = tile_first + (uint64_t)i * kEvalThreads + threadIdx.x;
        return out < n;
    }
    __device__ __forceinline__ bool full_tile(uint64_t &out) const
    {
        out = tile_first;
        return tile_first + kTile <= n;
    }
};

// ---------------------------------------------------------------------------------------------
// Lattice: vertex id -> digits -> per-axis coordinate table.
//   id <  ngrid : grid vertex, digits k_j base K+1, lattice index 2 k_j
//   id >= ngrid : centroid,    digits c_j base K,   lattice index 2 c_j + 1
// The table (d * (2K+1) doubles) is staged in shared memory when it fits.
// ---------------------------------------------------------------------------------------------
template <int P>
struct LatticeSrc {
    struct Params {
        const double *tab;
        uint64_t v0, nv, ngrid;
        int d, gen;
    };
    static constexpr int kP = P;
    static constexpr int kTile = kEvalThreads * P;
    static constexpr int kSmemTabMax = 4096;  // doubles
    const double *tab;
    uint64_t v0, nv, ngrid, tile_first;
    uint32_t rem[P];
    uint32_t isgrid;  // bit i set: point i is a grid vertex
    int d, gen, L1;

    static size_t smem_bytes(int d, int gen)
    {
        size_t e = (size_t)d * ((2u << gen) + 1);
        return e <= kSmemTabMax ? e * 8 : 0;
    }
    __device__ LatticeSrc(const Params &p, unsigned char *smem)
        : tab(p.tab), v0(p.v0), nv(p.nv), ngrid(p.ngrid), d(p.d), gen(p.gen)
    {
        L1 = (2 << gen) + 1;
        int e = d * L1;
        if (e <= kSmemTabMax) {
            double *s = reinterpret_cast<double *>(smem);
            for (int i = threadIdx.x; i < e; i += kEvalThreads) s[i] = __ldg(p.tab + i);
            tab = s;
            __syncthreads();
        }
    }
    __device__ uint64_t num_tiles() const { return (nv + kTile - 1) / kTile; }
    __device__ void begin_tile(uint64_t t)
    {
        tile_first = t * (uint64_t)kTile;
        isgrid = 0;
#pragma unroll
        for (int i = 0; i < P; ++i) {
            uint64_t id = v0 + tile_first + (uint64_t)i * kEvalThreads + threadIdx.x;
            bool g = id < ngrid;
            rem[i] = (uint32_t)(g ? id : id - ngrid);
            isgrid |= (uint32_t)g << i;
        }
    }
    __device__ __forceinline__ void load(int j, double (&x)[P])
    {
        const uint32_t K = 1u << gen;
        const double *row = tab + j * L1;
#pragma unroll
        for (int i = 0; i < P; ++i) {
            uint32_t p;
            if ((isgrid >> i) & 1u) {
                uint32_t q = rem[i] / (K + 1);
                p = 2 * (rem[i] - q * (K + 1));
                rem[i] = q;
            } else {
                p = 2 * (rem[i] & (K - 1)) + 1;
                rem[i] >>= gen;
            }
            x[i] = row[p];
        }
    }
    __device__ __forceinline__ bool valid(int i, uint64_t &out) const
    {
        out = tile_first + (uint64_t)i * kEvalThreads + threadIdx.x;
        return out < nv;
    }
    __device__ __forceinline__ bool full_tile(uint64_t &out) const
    {
        out = tile_first;
        return tile_first + kTile <= nv;
    }
};

}  // namespace shgo_b200
