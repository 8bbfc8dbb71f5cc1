// functors.cuh -- registered device objectives and constraint sets.
//
// Each functor states the numpy expression it replicates (the Python callable of the same name
// in shgo_b200/functors.py, i.e. what a reference user passes to shgo(func, ...), evaluated by
// VertexCacheField.compute_sfield, reference _vertex.py:338-352).  Unless a functor says
// otherwise the operation ORDER of that numpy expression is kept, every operation is an explicit
// round-to-nearest intrinsic (no FMA contraction; the TU is also built with -fmad=false), and
// np.sum is reproduced with numpy's own association order (sequential below 8 terms, 8-lane
// pairwise from 8 to 128 terms).  Polynomial / sqrt functors are therefore bit-identical to the
// reference's evaluation; cos/sin/exp go through CUDA's libdevice (<= 2 ulp from glibc/numpy),
// inside the 1e-12 acceptance tolerance.
//
// Interface:  F::P points per thread;  F::run(src, d, f[P], feas_mask) where bit i of feas_mask
// says whether point i satisfies every g >= 0 (tested as `g < 0` => infeasible, so NaN is
// feasible, reference _vertex.py:330-336).
#pragma once
#include "sources.cuh"

namespace shgo_b200 {

#define SHGO_ALL_P _Pragma("unroll") for (int i = 0; i < P; ++i)

// numpy add.reduce order over `nterms` per-axis terms; NS independent sums share one pass over
// the axes.  PAIR functors see (x_j, x_{j+1}) and have d-1 terms.
template <int P, int NS, bool PAIR, class Src, class Term>
__device__ __forceinline__ void np_sum_terms(Src &src, int nterms, Term term, double (&res)[NS][P])
{
    double xc[P], xn[P];
    int jload = 0;
    if (PAIR) {
        src.load(0, xn);
        jload = 1;
    }
    auto next = [&](double (&t)[NS][P]) {
        if (PAIR) {
            SHGO_ALL_P xc[i] = xn[i];
            src.load(jload++, xn);
        } else {
            src.load(jload++, xc);
        }
        SHGO_ALL_P
        {
            double tt[NS];
            term(xc[i], PAIR ? xn[i] : 0.0, tt);
#pragma unroll
            for (int s = 0; s < NS; ++s) t[s][i] = tt[s];
        }
    };
    const int nb = nterms >> 3;
    if (nb > 0) {
        double r[NS][8][P];
        for (int b = 0; b < nb; ++b) {
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                double t[NS][P];
                next(t);
#pragma unroll
                for (int s = 0; s < NS; ++s)
                    SHGO_ALL_P r[s][jj][i] = b == 0 ? t[s][i] : __dadd_rn(r[s][jj][i], t[s][i]);
            }
        }
#pragma unroll
        for (int s = 0; s < NS; ++s)
            SHGO_ALL_P
            {
                double a = __dadd_rn(__dadd_rn(r[s][0][i], r[s][1][i]),
                                     __dadd_rn(r[s][2][i], r[s][3][i]));
                double b2 = __dadd_rn(__dadd_rn(r[s][4][i], r[s][5][i]),
                                      __dadd_rn(r[s][6][i], r[s][7][i]));
                res[s][i] = __dadd_rn(a, b2);
            }
    } else {
#pragma unroll
        for (int s = 0; s < NS; ++s) SHGO_ALL_P res[s][i] = 0.0;
    }
    for (int j = nb * 8; j < nterms; ++j) {
        double t[NS][P];
        next(t);
#pragma unroll
        for (int s = 0; s < NS; ++s) SHGO_ALL_P res[s][i] = __dadd_rn(res[s][i], t[s][i]);
    }
}

constexpr double kTwoPi = 6.283185307179586;  // python: 2*np.pi
constexpr double kE = 2.718281828459045;      // np.e

// scipy.optimize.rosen:  sum(100.0*(x[1:]-x[:-1]**2.0)**2.0 + (1-x[:-1])**2.0, axis=0)
struct Rosenbrock {
    static constexpr int P = 4;
    static constexpr int kMinBlocks = 1;  // resident CTAs per SM the register budget must allow
    static bool dim_ok(int d) { return d >= 2; }
    template <class Src>
    static __device__ __forceinline__ void run(Src &src, int d, double (&f)[P], uint32_t &feas)
    {
        double s[1][P];
        np_sum_terms<P, 1, true>(
            src, d - 1,
            [](double x, double xn, double (&t)[1]) {
                double a = __dmul_rn(x, x);
                double b = __dsub_rn(xn, a);
                double c = __dmul_rn(b, b);
                double e = __dmul_rn(100.0, c);
                double g = __dsub_rn(1.0, x);
                double h = __dmul_rn(g, g);
                t[0] = __dadd_rn(e, h);
            },
            s);
        SHGO_ALL_P f[i] = s[0][i];
        feas = ~0u;
    }
};

// 10*len(x) + np.sum(x**2 - 10*np.cos(2*np.pi*x))
struct Rastrigin {
    static constexpr int P = 4;
    static constexpr int kMinBlocks = 3;  // resident CTAs per SM the register budget must allow
    static bool dim_ok(int d) { return d >= 1; }
    template <class Src>
    static __device__ __forceinline__ void run(Src &src, int d, double (&f)[P], uint32_t &feas)
    {
        double s[1][P];
        np_sum_terms<P, 1, false>(
            src, d,
            [](double x, double, double (&t)[1]) {
                double a = __dmul_rn(x, x);
                double c = cos(__dmul_rn(kTwoPi, x));
                t[0] = __dsub_rn(a, __dmul_rn(10.0, c));
            },
            s);
        const double base = (double)(10 * d);
        SHGO_ALL_P f[i] = __dadd_rn(base, s[0][i]);
        feas = ~0u;
    }
};

// -20*np.exp(-0.2*np.sqrt(np.sum(x**2)/d)) - np.exp(np.sum(np.cos(2*np.pi*x))/d) + 20 + np.e
struct Ackley {
    static constexpr int P = 2;
    static constexpr int kMinBlocks = 3;  // resident CTAs per SM the register budget must allow
    static bool dim_ok(int d) { return d >= 1; }
    template <class Src>
    static __device__ __forceinline__ void run(Src &src, int d, double (&f)[P], uint32_t &feas)
    {
        double s[2][P];
        np_sum_terms<P, 2, false>(
            src, d,
            [](double x, double, double (&t)[2]) {
                t[0] = __dmul_rn(x, x);
                t[1] = cos(__dmul_rn(kTwoPi, x));
            },
            s);
        const double dd = (double)d;
        SHGO_ALL_P
        {
            double A = __dmul_rn(-20.0, exp(__dmul_rn(-0.2, __dsqrt_rn(__ddiv_rn(s[0][i], dd)))));
            double B = __dsub_rn(A, exp(__ddiv_rn(s[1][i], dd)));
            f[i] = __dadd_rn(__dadd_rn(B, 20.0), kE);
        }
        feas = ~0u;
    }
};

// 0.5*np.sum(x**4 - 16*x**2 + 5*x).  Not defined by the reference; the
// registered device form is the Horner/FMA chain  acc = fma(fma(fma(x,x,-16),x,5),x,acc)  over
// the axes in order, f = 0.5*acc -- 3 DFMA per axis.  It agrees with the numpy expression to
// ~1e-15 relative to the term magnitudes (tests pin 1e-12) and bit-for-bit with the oracle's C
// restatement of the same chain.
struct StyblinskiTang {
    static constexpr int P = 8;
    static constexpr int kMinBlocks = 1;  // resident CTAs per SM the register budget must allow
    static bool dim_ok(int d) { return d >= 1; }
    template <class Src>
    static __device__ __forceinline__ void run(Src &src, int d, double (&f)[P], uint32_t &feas)
    {
        double acc[P];
        SHGO_ALL_P acc[i] = 0.0;
        for (int j = 0; j < d; ++j) {
            double x[P];
            src.load(j, x);
            SHGO_ALL_P
            {
                double q = __fma_rn(x[i], x[i], -16.0);
                q = __fma_rn(q, x[i], 5.0);
                acc[i] = __fma_rn(q, x[i], acc[i]);
            }
        }
        SHGO_ALL_P f[i] = __dmul_rn(0.5, acc[i]);
        feas = ~0u;
    }
};

// reference _shgo.py:329-333
struct Eggholder {
    static constexpr int P = 2;
    static constexpr int kMinBlocks = 3;  // resident CTAs per SM the register budget must allow
    static bool dim_ok(int d) { return d == 2; }
    template <class Src>
    static __device__ __forceinline__ void run(Src &src, int, double (&f)[P], uint32_t &feas)
    {
        double x0[P], x1[P];
        src.load(0, x0);
        src.load(1, x1);
        SHGO_ALL_P
        {
            double y47 = __dadd_rn(x1[i], 47.0);
            double a = __dmul_rn(-y47, sin(__dsqrt_rn(fabs(__dadd_rn(__ddiv_rn(x0[i], 2.0), y47)))));
            double b = __dmul_rn(x0[i], sin(__dsqrt_rn(fabs(__dsub_rn(x0[i], y47)))));
            f[i] = __dsub_rn(a, b);
        }
        feas = ~0u;
    }
};

// reference _shgo.py:408-418 (cattle-feed, HS73): objective and, when CONS, g1 and g2
template <bool CONS>
struct CattleFeed {
    static constexpr int P = 4;
    static constexpr int kMinBlocks = 1;
    static bool dim_ok(int d) { return d == 4; }
    template <class Src>
    static __device__ __forceinline__ void run(Src &src, int, double (&f)[P], uint32_t &feas)
    {
        double x[4][P];
#pragma unroll
        for (int j = 0; j < 4; ++j) src.load(j, x[j]);
        feas = ~0u;
        SHGO_ALL_P
        {
            const double a = x[0][i], b = x[1][i], c = x[2][i], e = x[3][i];
            f[i] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(24.55, a), __dmul_rn(26.75, b)),
                                       __dmul_rn(39.0, c)),
                             __dmul_rn(40.50, e));
            if (CONS) {
                double g1 = __dsub_rn(
                    __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(2.3, a), __dmul_rn(5.6, b)),
                                        __dmul_rn(11.1, c)),
                              __dmul_rn(1.3, e)),
                    5.0);
                double lin = __dsub_rn(
                    __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(12.0, a), __dmul_rn(11.9, b)),
                                        __dmul_rn(41.8, c)),
                              __dmul_rn(52.1, e)),
                    21.0);
                double q = __dadd_rn(
                    __dadd_rn(__dadd_rn(__dmul_rn(0.28, __dmul_rn(a, a)),
                                        __dmul_rn(0.19, __dmul_rn(b, b))),
                              __dmul_rn(20.5, __dmul_rn(c, c))),
                    __dmul_rn(0.62, __dmul_rn(e, e)));
                double g2 = __dsub_rn(lin, __dmul_rn(1.645, __dsqrt_rn(q)));
                if (g1 < 0.0 || g2 < 0.0) feas &= ~(1u << i);
            }
        }
    }
};

// x[0]**2 + x[1]**2 + ... left to right (reference tests/test__shgo.py:41-42,242-246); when
// CONS the constraint -(numpy.sum(x, axis=0) - 6.0) >= 0 of tests/test__shgo.py:44-45 (d < 8,
// where numpy's sum is sequential from 0.0)
template <bool CONS>
struct Sphere {
    static constexpr int P = 4;
    static constexpr int kMinBlocks = 1;
    static bool dim_ok(int d) { return d >= 1 && (!CONS || d < 8); }
    template <class Src>
    static __device__ __forceinline__ void run(Src &src, int d, double (&f)[P], uint32_t &feas)
    {
        double acc[P], sx[P];
        SHGO_ALL_P sx[i] = 0.0;
        for (int j = 0; j < d; ++j) {
            double x[P];
            src.load(j, x);
            SHGO_ALL_P
            {
                double sq = __dmul_rn(x[i], x[i]);
                acc[i] = j == 0 ? sq : __dadd_rn(acc[i], sq);
                if (CONS) sx[i] = __dadd_rn(sx[i], x[i]);
            }
        }
        feas = ~0u;
        SHGO_ALL_P
        {
            f[i] = acc[i];
            if (CONS) {
                double g = -__dsub_rn(sx[i], 6.0);
                if (g < 0.0) feas &= ~(1u << i);
            }
        }
    }
};

}  // namespace shgo_b200
