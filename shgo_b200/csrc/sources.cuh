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
// The P points of a thread are visited in Gray-code order of i (0,1,3,2,6,7,5,4), so that each
// one costs ONE XOR with w_j[0], w_j[1] or w_j[2] (registers) -- no per-thread generator state and
// no table look-up per point.  Per (tile, axis) a thread reads LOW, AH and one 40-byte record
// {rs, cc, lb, w0, w1, w2} from shared memory (broadcasts): 0.5 LDS per point-coordinate at P = 8
// (round 1 read AI and the three constants per point: 2.25).
// AH changes per tile.  It is computed for kRing tiles at a time by kRing*d threads into one half of
// a double buffer WHILE the block evaluates the other half: one __syncthreads per kRing tiles
// (round 1: two per tile, 29 % of the kernel's stall samples).
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
    struct __align__(16) AxisRec {  // 48 bytes, three 16-byte broadcast loads
        double rs, cc;  // cc = -2^52 rs; when the axis is `exact` (see below): cc = lb - 2^52 rs
        double lb;
        uint32_t w0, w1;
        uint32_t w2, pad0, pad1, pad2;
    };
    static constexpr int kP = P;
    static constexpr int kTile = kEvalThreads * P;
    static constexpr int kRing = 4;  // tiles whose AH words are computed together
    static_assert((P & (P - 1)) == 0 && P <= 8, "P must be a power of two <= 8");

    uint32_t *low, *whi, *ah;  // ah: [2][kRing][d]
    const uint32_t *ah_cur;    // the current tile's d words
    AxisRec *rec;
    uint64_t k0, n, kb, tile_first;
    int d;
    uint32_t calls;
    // One rounding instead of two: when u * rs is exact for every 30-bit u (rs has at most 23
    // significant bits) and lb - 2^52 rs is representable, fl(fl(u rs) + lb) = fl(u rs + lb) =
    // fma(2^52 + u, rs, lb - 2^52 rs) -- still bit-identical to numpy's multiply-then-add, with one
    // FP64 instruction per coordinate (e.g. bounds (-5, 5), (0, 2), (0, 1); not (-5.12, 5.12)).
    // Decided on the device from lb / ub, for all axes together.
    bool exact;

    static size_t smem_bytes(int d)
    {
        size_t words = (size_t)d * (256 + 22 + 2 * kRing);
        words = (words + 3) & ~(size_t)3;
        return words * 4 + (size_t)d * sizeof(AxisRec) + 8 + (size_t)d * 8;
    }

    __device__ __forceinline__ void compute_ah(uint64_t first_tile, uint64_t stride, uint32_t *dst)
    {
        // threads q*d + j, q < kRing: AH_j of tile first_tile + q*stride
        const int t = threadIdx.x;
        if (t < kRing * d) {
            const int q = t / d, j = t - q * d;
            const uint64_t tf = kb + (first_tile + (uint64_t)q * stride) * (uint64_t)kTile;
            uint32_t M = (uint32_t)(tf >> 8), a = 0;
            for (int c = 0; M != 0; ++c, M >>= 1)
                if (M & 1u) a ^= whi[j * 22 + c];
            dst[q * d + j] = a;
        }
    }

    __device__ SobolSrc(const Params &p, unsigned char *smem) : k0(p.k0), n(p.n), d(p.d), calls(0)
    {
        low = reinterpret_cast<uint32_t *>(smem);
        whi = low + d * 256;
        ah = whi + d * 22;
        size_t words = ((size_t)d * (256 + 22 + 2 * kRing) + 3) & ~(size_t)3;
        rec = reinterpret_cast<AxisRec *>(smem + words * 4);
        kb = k0 & ~(uint64_t)(kTile - 1);
        const int t = threadIdx.x;
        if (t == 0) *not_exact_flag() = 1;
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
        __syncthreads();
        if (t < d) {
            double r = __dsub_rn(__ldg(p.ub + t), __ldg(p.lb + t));
            double s = r * 9.313225746154785e-10;  // 2^-30, exact scaling
            AxisRec a;
            a.rs = s;
            a.cc = -(s * 4503599627370496.0);  // -2^52 * rs, exact
            a.lb = __ldg(p.lb + t);
            {
                // u*rs exact <=> the low 30 bits of rs's 53-bit significand are zero (or rs == 0)
                const unsigned long long mant = (unsigned long long)__double_as_longlong(s) & 0xfffffffffffffull;
                bool ok = s == 0.0 || ((mant & 0x3fffffffull) == 0 && fabs(s) > 1e-290);
                // c1 = lb + cc exact?  (TwoSum error term)
                const double c1 = __dadd_rn(a.lb, a.cc);
                const double bb = __dsub_rn(c1, a.lb);
                const double err = __dadd_rn(__dsub_rn(a.lb, __dsub_rn(c1, bb)), __dsub_rn(a.cc, bb));
                ok = ok && err == 0.0 && isfinite(c1);
                if (!ok) atomicAnd(not_exact_flag(), 0);
                rec_c1(t) = c1;
            }
            a.w0 = whi[t * 22], a.w1 = whi[t * 22 + 1], a.w2 = whi[t * 22 + 2];
            a.pad0 = a.pad1 = a.pad2 = 0;
            rec[t] = a;
        }
        // AH of this block's first kRing tiles; begin_tile() keeps one group ahead from here on
        compute_ah(blockIdx.x, gridDim.x, ah);
        __syncthreads();
        exact = *not_exact_flag() != 0;
        __syncthreads();
        if (exact && t < d) rec[t].cc = rec_c1(t);
        __syncthreads();
    }
    // scratch behind the records: one flag word + d doubles (only used inside the constructor)
    __device__ __forceinline__ int *not_exact_flag() { return reinterpret_cast<int *>(rec + d); }
    __device__ __forceinline__ double &rec_c1(int j) { return reinterpret_cast<double *>(rec + d)[1 + j]; }

    __device__ uint64_t num_tiles() const { return (k0 + n - kb + kTile - 1) / kTile; }

    // `t` must advance by gridDim.x from blockIdx.x (the persistent loop of the evaluation kernels)
    __device__ void begin_tile(uint64_t t)
    {
        const uint32_t q = calls % kRing, grp = calls / kRing;
        if (q == 0) {
            // group `grp` is complete (it was computed while group grp-1 was evaluated) and nobody
            // reads group grp-1 any more: its half of the buffer takes group grp+1
            __syncthreads();
            compute_ah(t + (uint64_t)kRing * gridDim.x, gridDim.x, ah + ((grp + 1) & 1u) * kRing * d);
        }
        ah_cur = ah + ((grp & 1u) * kRing + q) * d;
        tile_first = kb + t * (uint64_t)kTile;
        ++calls;
    }

    __device__ __forceinline__ void load(int j, double (&x)[P])
    {
        if (exact) load_t<true>(j, x); else load_t<false>(j, x);
    }
    // The evaluation kernels pick the path once per launch and run their tile loop on a View, so
    // that the choice is a template parameter inside the per-axis loop (a runtime flag there is
    // if-converted into DADD + 2 FSEL per coordinate).
    static constexpr bool kHasFastPath = true;
    template <bool EX>
    struct View {
        SobolSrc &s;
        __device__ __forceinline__ void begin_tile(uint64_t t) { s.begin_tile(t); }
        __device__ __forceinline__ void load(int j, double (&x)[P]) { s.template load_t<EX>(j, x); }
        __device__ __forceinline__ bool valid(int i, uint64_t &out) const { return s.valid(i, out); }
        __device__ __forceinline__ bool full_tile(uint64_t &out) const { return s.full_tile(out); }
    };
    __device__ __forceinline__ bool fast_path() const { return exact; }

    template <bool EX>
    __device__ __forceinline__ void load_t(int j, double (&x)[P])
    {
        // D = 2^52 + u as a bit pattern; the Gray-code steps flip bits of its low word in place
        unsigned long long dbits = 0x4330000000000000ull | (low[j * 256 + threadIdx.x] ^ ah_cur[j]);
        const AxisRec a = rec[j];
        const uint32_t w[3] = {a.w0, a.w1, a.w2};
        int i = 0;
#pragma unroll
        for (int g = 0; g < P; ++g) {
            if (g > 0) {  // Gray-code step: flip bit ctz(g) of i
                const int bit = (g & 1) ? 0 : ((g & 2) ? 1 : 2);
                i ^= 1 << bit;
                dbits ^= (unsigned long long)w[bit];
            }
            const double t = __fma_rn(__longlong_as_double((long long)dbits), a.rs, a.cc);
            x[i] = EX ? t : __dadd_rn(t, a.lb);
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

    static constexpr bool kHasFastPath = false;
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
    static constexpr bool kHasFastPath = false;
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
