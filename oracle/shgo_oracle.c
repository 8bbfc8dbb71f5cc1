/*
 * shgo_oracle.c -- CPU ORACLE for the SHGO sampling-and-classification hot path.
 *
 * THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, the smoke check in
 * __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product (shgo_b200/) never imports, links or executes anything under oracle/.
 *
 * It restates, in plain scalar C, what the reference (Stefan-Endres/shgo, pure Python on
 * top of scipy/numpy) computes on the path
 *     generate points -> evaluate f, g -> mask infeasible -> star test (minimiser pool).
 * Every function cites the reference file:line it follows (paths relative to the
 * reference checkout; "scipy/" = the installed scipy 1.18.1).
 *
 * Parity pinning: the reference holds NO golden vectors for this path (its tests are
 * end-to-end).  This oracle is therefore pinned against outputs of the
 * reference itself, generated in the build container by tests/golden/make_golden.py
 * (imports /root/reference unmodified) and committed under tests/golden/*.npz:
 * Sobol points bit-equal, edge sets equal, f within 1e-12 (bit-equal for the polynomial
 * objectives), feasibility and minimiser-pool index sets equal.  tests/test_oracle_*.py
 * run those checks on CPU.
 *
 * Third-party arithmetic on the path whose source is not in the reference checkout:
 *   - scipy 1.18.1 scipy.stats._sobol (Cython, binary only): restated from the published
 *     Joe-Kuo / Bratley-Fox algorithm, anchored on _shgo.py:703-711,1441-1444 and checked
 *     bit-for-bit against the installed scipy.
 *   - qhull (scipy.spatial.Delaunay): NOT restated; both sides consume the same
 *     Tri.simplices array.
 *   - numpy ufuncs (cos, exp, sqrt, power): libm here; agreement 1e-12, not bit-equal.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off; FMA contraction would break the
 * two-rounding `u*(ub-lb)+lb` of _shgo.py:1446-1449).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define SOBOL_BITS 30

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline is meant to use the host */
void oracle_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------ */
/* a1. Sobol direction numbers and points                                                */
/* ------------------------------------------------------------------------------------ */

/*
 * Direction matrix sv[d][30] from Joe-Kuo rows (dimension j>=2: degree s, interior
 * coefficients a, initial m_1..m_s).  Follows what scipy's _initialize_v produces
 * (call site scipy/stats/_qmc.py:1793-1794; reference call site _shgo.py:703-704);
 * the table itself is the one the reference ships as shgo/_shgo_lib/sobol_vec.gz.
 * m is [d][18] (row 0 unused: dimension 1 has v_i = 2^(30-i)).
 */
void oracle_sobol_dirnums(int d, const int32_t *s_deg, const int32_t *a_coef,
                          const int32_t *m_init, uint32_t *sv)
{
    for (int j = 0; j < d; ++j) {
        uint32_t *v = sv + (size_t)j * SOBOL_BITS;
        if (j == 0) {
            for (int i = 0; i < SOBOL_BITS; ++i) v[i] = 1u << (SOBOL_BITS - 1 - i);
            continue;
        }
        int s = s_deg[j];
        uint32_t a = (uint32_t)a_coef[j];
        const int32_t *m = m_init + (size_t)j * 18;
        for (int i = 0; i < s && i < SOBOL_BITS; ++i)
            v[i] = (uint32_t)m[i] << (SOBOL_BITS - 1 - i);
        for (int i = s; i < SOBOL_BITS; ++i) {
            uint32_t x = v[i - s] ^ (v[i - s] >> s);
            for (int k = 1; k < s; ++k)
                if ((a >> (s - 1 - k)) & 1u) x ^= v[i - k];
            v[i] = x;
        }
    }
}

static int ctz64(uint64_t k)
{
    int c = 0;
    while (!(k & 1u)) { k >>= 1; ++c; }
    return c;
}

/*
 * Points k0 .. k0+n-1 of the unscrambled sequence, scaled to the bounds, written row-major
 * (n,d) like the reference's self.C.
 *   - sequential Gray-code update exactly as scipy's _draw does it
 *     (scipy/stats/_qmc.py:1862-1884: point_k = point_{k-1} XOR sv[:, index of the lowest
 *     zero bit of k-1]); first point is the zero vector (_qmc.py:1803-1811);
 *   - k0 > 0 replays the same recurrence like _fast_forward (engine state persists across
 *     shgo iterations, _shgo.py:1441-1444);
 *   - scaling is the UNFUSED multiply-then-add of _shgo.py:1446-1449.
 */
void oracle_sobol_points(const uint32_t *sv, int d, uint64_t k0, uint64_t n,
                         const double *lb, const double *ub, double *out)
{
    uint32_t *q = (uint32_t *)calloc((size_t)d, sizeof(uint32_t));
    for (uint64_t k = 1; k <= k0; ++k) {
        int c = ctz64(k);
        for (int j = 0; j < d; ++j) q[j] ^= sv[(size_t)j * SOBOL_BITS + c];
    }
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t k = k0 + i;
        if (i > 0) {
            int c = ctz64(k);
            for (int j = 0; j < d; ++j) q[j] ^= sv[(size_t)j * SOBOL_BITS + c];
        }
        for (int j = 0; j < d; ++j) {
            double u = (double)q[j] * (1.0 / 1073741824.0); /* 2^-30, exact */
            double r = ub[j] - lb[j];
            double t = u * r;
            out[i * (uint64_t)d + j] = t + lb[j];
        }
    }
    free(q);
}

/* ------------------------------------------------------------------------------------ */
/* numpy's add.reduce association order for a contiguous 1-D double array               */
/* (numpy/_core/src/umath/loops_utils.h.src pairwise sum; probed in SURVEY.md section 7)        */
/* ------------------------------------------------------------------------------------ */
static double np_sum(const double *a, int n)
{
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r += a[i];
        return r;
    }
    if (n <= 128) {
        double r[8];
        int i;
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return np_sum(a, n2) + np_sum(a + n2, n - n2);
}

/* ------------------------------------------------------------------------------------ */
/* a3. objectives (what VertexCacheField.compute_sfield calls, _vertex.py:338-352)       */
/* The op order below is the canonical numpy expression of each objective, which is what */
/* shgo_b200/functors.py exposes to users and what the device functors replicate.        */
/* ------------------------------------------------------------------------------------ */

#define ORACLE_MAXD 128

enum {
    F_ROSENBROCK = 0,
    F_RASTRIGIN = 1,
    F_ACKLEY = 2,
    F_STYBLINSKI_TANG = 3,
    F_EGGHOLDER = 4,
    F_CATTLE_FEED = 5,
    F_SPHERE = 6
};

/* scipy.optimize.rosen (scipy/optimize/_optimize.py):
 *   sum(100.0*(x[1:]-x[:-1]**2.0)**2.0 + (1-x[:-1])**2.0, axis=0)                       */
static double f_rosenbrock(const double *x, int d)
{
    double t[ORACLE_MAXD];
    for (int i = 0; i < d - 1; ++i) {
        double a = x[i] * x[i];
        double b = x[i + 1] - a;
        double c = b * b;
        double e = 100.0 * c;
        double g = 1.0 - x[i];
        double h = g * g;
        t[i] = e + h;
    }
    return np_sum(t, d - 1);
}

/* 10*d + np.sum(x**2 - 10*np.cos(2*np.pi*x)) */
static double f_rastrigin(const double *x, int d)
{
    double t[ORACLE_MAXD];
    const double two_pi = 2 * M_PI;
    for (int i = 0; i < d; ++i) {
        double a = x[i] * x[i];
        double c = cos(two_pi * x[i]);
        double e = 10.0 * c;
        t[i] = a - e;
    }
    return (double)(10 * d) + np_sum(t, d);
}

/* -20*np.exp(-0.2*np.sqrt(np.sum(x**2)/d)) - np.exp(np.sum(np.cos(2*np.pi*x))/d) + 20 + np.e */
static double f_ackley(const double *x, int d)
{
    double t[ORACLE_MAXD], c[ORACLE_MAXD];
    const double two_pi = 2 * M_PI;
    for (int i = 0; i < d; ++i) {
        t[i] = x[i] * x[i];
        c[i] = cos(two_pi * x[i]);
    }
    double s1 = np_sum(t, d);
    double s2 = np_sum(c, d);
    double A = -20.0 * exp(-0.2 * sqrt(s1 / (double)d));
    double B = A - exp(s2 / (double)d);
    double C = B + 20.0;
    return C + M_E;
}

/* 0.5*np.sum(x**4 - 16*x**2 + 5*x)   (not defined in the reference; conventional form,
 * SURVEY.md section 8c golden table).  The registered form is the Horner chain
 *   acc = fma(fma(fma(x,x,-16),x,5),x,acc)  over the axes in order, f = 0.5*acc
 * (C99 fma = one rounding, identical to the device DFMA).  It agrees with numpy's evaluation of
 * the expression above (x**4 goes through pow()) to ~1e-15 of the term magnitudes -- inside the
 * 1e-12 tolerance of the acceptance criteria; tests/test_oracle_golden.py checks that and that
 * the minimiser pools of the golden cases are unchanged. */
static double f_styblinski_tang(const double *x, int d)
{
    double acc = 0.0;
    for (int i = 0; i < d; ++i) {
        double q = fma(x[i], x[i], -16.0);
        q = fma(q, x[i], 5.0);
        acc = fma(q, x[i], acc);
    }
    return 0.5 * acc;
}

/* _shgo.py:329-333 (also tests/test__shgo.py:181-186) */
static double f_eggholder(const double *x, int d)
{
    (void)d;
    double y47 = x[1] + 47.0;
    double a = -y47 * sin(sqrt(fabs(x[0] / 2.0 + y47)));
    double b = x[0] * sin(sqrt(fabs(x[0] - y47)));
    return a - b;
}

/* cattle-feed objective, _shgo.py:408-409 */
static double f_cattle_feed(const double *x, int d)
{
    (void)d;
    return ((24.55 * x[0] + 26.75 * x[1]) + 39.0 * x[2]) + 40.50 * x[3];
}

/* x[0]**2 + x[1]**2 + ...  left to right (tests/test__shgo.py:41-42, :242-246) */
static double f_sphere(const double *x, int d)
{
    double r = x[0] * x[0];
    for (int i = 1; i < d; ++i) r = r + x[i] * x[i];
    return r;
}

typedef double (*objective_fn)(const double *, int);

static objective_fn pick_objective(int id)
{
    switch (id) {
    case F_ROSENBROCK: return f_rosenbrock;
    case F_RASTRIGIN: return f_rastrigin;
    case F_ACKLEY: return f_ackley;
    case F_STYBLINSKI_TANG: return f_styblinski_tang;
    case F_EGGHOLDER: return f_eggholder;
    case F_CATTLE_FEED: return f_cattle_feed;
    case F_SPHERE: return f_sphere;
    default: return NULL;
    }
}

/* ------------------------------------------------------------------------------------ */
/* a4. constraint sets: feasible <=> every g_i(x) >= 0, tested as `g(x) < 0.0` so NaN is  */
/* feasible (_vertex.py:330-336)                                                          */
/* ------------------------------------------------------------------------------------ */
enum { G_NONE = 0, G_CATTLE_FEED = 1, G_SUM_LE_6 = 2 };

/* _shgo.py:411-418 */
static double g_cattle1(const double *x)
{
    return (((2.3 * x[0] + 5.6 * x[1]) + 11.1 * x[2]) + 1.3 * x[3]) - 5.0;
}
static double g_cattle2(const double *x)
{
    double lin = ((((12.0 * x[0] + 11.9 * x[1]) + 41.8 * x[2]) + 52.1 * x[3]) - 21.0);
    double q = ((0.28 * (x[0] * x[0]) + 0.19 * (x[1] * x[1])) + 20.5 * (x[2] * x[2])) +
               0.62 * (x[3] * x[3]);
    return lin - 1.645 * sqrt(q);
}

static int feasible(int gset, const double *x, int d)
{
    switch (gset) {
    case G_NONE: return 1;
    case G_CATTLE_FEED:
        if (g_cattle1(x) < 0.0) return 0;
        if (g_cattle2(x) < 0.0) return 0;
        return 1;
    case G_SUM_LE_6: { /* tests/test__shgo.py:44-45: -(numpy.sum(x, axis=0) - 6.0) */
        double g = -(np_sum(x, d) - 6.0);
        return !(g < 0.0);
    }
    default: return 1;
    }
}

/*
 * Evaluate objective + constraints on (n,d) row-major points.
 * f[i] = +inf if infeasible (_vertex.py:334) or NaN (_vertex.py:351-352); feas[i] in {0,1}.
 * Returns -1 for an unknown functor.
 */
int oracle_eval(int functor, int gset, int d, uint64_t n, const double *x, double *f,
                uint8_t *feas)
{
    objective_fn fn = pick_objective(functor);
    if (!fn || d > ORACLE_MAXD) return -1;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        const double *xi = x + (uint64_t)i * d;
        int ok = feasible(gset, xi, d);
        double v = INFINITY;
        if (ok) {
            v = fn(xi, d);
            if (isnan(v)) v = INFINITY;
        }
        f[i] = v;
        if (feas) feas[i] = (uint8_t)ok;
    }
    return 0;
}

/* Fused generate + evaluate, shard-parallel: the CPU counterpart of the device hot path. */
int oracle_sobol_eval(const uint32_t *sv, int functor, int gset, int d, uint64_t k0, uint64_t n,
                      const double *lb, const double *ub, double *f, uint8_t *feas)
{
    objective_fn fn = pick_objective(functor);
    if (!fn || d > ORACLE_MAXD) return -1;
    const uint64_t chunk = 1u << 14;
    int64_t nchunks = (int64_t)((n + chunk - 1) / chunk);
#pragma omp parallel
    {
        double *buf = (double *)malloc(chunk * (size_t)d * sizeof(double));
#pragma omp for schedule(static)
        for (int64_t c = 0; c < nchunks; ++c) {
            uint64_t i0 = (uint64_t)c * chunk;
            uint64_t m = n - i0 < chunk ? n - i0 : chunk;
            /* closed-form state at k0+i0 would do; the oracle keeps the replayed recurrence
             * but starts it from the Gray-code state to stay O(n) */
            uint64_t k = k0 + i0;
            uint64_t gray = k ^ (k >> 1);
            uint32_t q[ORACLE_MAXD];
            for (int j = 0; j < d; ++j) {
                uint32_t v = 0;
                for (int b = 0; b < SOBOL_BITS; ++b)
                    if ((gray >> b) & 1u) v ^= sv[(size_t)j * SOBOL_BITS + b];
                q[j] = v;
            }
            for (uint64_t i = 0; i < m; ++i) {
                if (i > 0) {
                    int cz = ctz64(k + i);
                    for (int j = 0; j < d; ++j) q[j] ^= sv[(size_t)j * SOBOL_BITS + cz];
                }
                for (int j = 0; j < d; ++j) {
                    double u = (double)q[j] * (1.0 / 1073741824.0);
                    double t = u * (ub[j] - lb[j]);
                    buf[i * d + j] = t + lb[j];
                }
            }
            for (uint64_t i = 0; i < m; ++i) {
                const double *xi = buf + i * d;
                int ok = feasible(gset, xi, d);
                double v = INFINITY;
                if (ok) {
                    v = fn(xi, d);
                    if (isnan(v)) v = INFINITY;
                }
                f[i0 + i] = v;
                if (feas) feas[i0 + i] = (uint8_t)ok;
            }
        }
        free(buf);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* a5. vf_to_vv: vertex-face mesh -> vertex-vertex graph (_complex.py:1053-1074)          */
/* dim > 1: itertools.combinations(s, dim) connects only e[0]-e[1] of every combination,  */
/* i.e. the three edges among simplex slots 0,1,2.  dim == 1: the   */
/* pair itself.  Output: symmetric CSR with sorted, de-duplicated rows; rows of vertices  */
/* that never occupy a connecting slot are empty (= vertex absent from HC.V).             */
/* ------------------------------------------------------------------------------------ */
static int cmp_u64(const void *a, const void *b)
{
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return (x > y) - (x < y);
}

/* returns nnz; col must hold 6*S entries (2*S when dp1 == 2) */
int64_t oracle_vf_to_vv(const int32_t *simplices, uint64_t S, int dp1, uint64_t n,
                        int64_t *row_ptr, int32_t *col)
{
    int nslot = dp1 >= 3 ? 3 : 2;
    int epairs = nslot == 3 ? 3 : 1;
    uint64_t cap = S * (uint64_t)epairs * 2u;
    uint64_t *keys = (uint64_t *)malloc((cap ? cap : 1) * sizeof(uint64_t));
    uint64_t m = 0;
    static const int pa[3] = {0, 0, 1}, pb[3] = {1, 2, 2};
    for (uint64_t s = 0; s < S; ++s) {
        const int32_t *sp = simplices + s * (uint64_t)dp1;
        for (int e = 0; e < epairs; ++e) {
            uint32_t u = (uint32_t)sp[pa[e]], v = (uint32_t)sp[pb[e]];
            if (u == v) continue; /* connect(): `v is not self` (_vertex.py:123) */
            keys[m++] = ((uint64_t)u << 32) | v;
            keys[m++] = ((uint64_t)v << 32) | u;
        }
    }
    qsort(keys, m, sizeof(uint64_t), cmp_u64);
    memset(row_ptr, 0, (n + 1) * sizeof(int64_t));
    int64_t nnz = 0;
    for (uint64_t i = 0; i < m; ++i) {
        if (i > 0 && keys[i] == keys[i - 1]) continue;
        uint32_t u = (uint32_t)(keys[i] >> 32);
        col[nnz++] = (int32_t)(keys[i] & 0xffffffffu);
        row_ptr[u + 1]++;
    }
    for (uint64_t i = 0; i < n; ++i) row_ptr[i + 1] += row_ptr[i];
    free(keys);
    return nnz;
}

/* ------------------------------------------------------------------------------------ */
/* a6. star test: is_min[v] = all(f[v] < f[w] for w in nn[v]) (_vertex.py:144-151), strict; */
/* a vertex with no neighbours is absent from the complex (fact 5) and reports 0.         */
/* ------------------------------------------------------------------------------------ */
void oracle_star_csr(const int64_t *row_ptr, const int32_t *col, const double *f, uint64_t row0,
                     uint64_t rows, uint8_t *is_min)
{
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < (int64_t)rows; ++r) {
        uint64_t v = row0 + (uint64_t)r;
        int64_t b = row_ptr[v], e = row_ptr[v + 1];
        int ok = e > b;
        double fv = f[v];
        for (int64_t k = b; k < e && ok; ++k)
            if (!(fv < f[col[k]])) ok = 0;
        is_min[r] = (uint8_t)ok;
    }
}

/* minimizers(): indices of pool members in ascending vertex order (_shgo.py:1109-1157) */
uint64_t oracle_compact_pool(const uint8_t *is_min, uint64_t n, int64_t *idx)
{
    uint64_t c = 0;
    for (uint64_t i = 0; i < n; ++i)
        if (is_min[i]) idx[c++] = (int64_t)i;
    return c;
}

/* a11. find_lowest_vertex (_shgo.py:863-880): first strict minimum; -1 if everything is inf */
int64_t oracle_argmin(const double *f, uint64_t n, const uint8_t *present)
{
    int64_t best = -1;
    double fb = INFINITY;
    for (uint64_t i = 0; i < n; ++i) {
        if (present && !present[i]) continue;
        if (f[i] < fb) { fb = f[i]; best = (int64_t)i; }
    }
    return best;
}

/* nfev = feasible vertices actually present in the complex (_vertex.py:184,347) */
uint64_t oracle_nfev(const int64_t *row_ptr, const uint8_t *feas, uint64_t n)
{
    uint64_t c = 0;
    for (uint64_t i = 0; i < n; ++i)
        if (row_ptr[i + 1] > row_ptr[i] && (!feas || feas[i])) ++c;
    return c;
}

/* f2. local bounds of a pool vertex, SHGO.construct_lcb_simplicial (_shgo.py:1246-1270):      */
/* per axis the neighbour coordinate nearest below / above the vertex's own, starting from   */
/* the global bounds.  x is row-major [n][d]; lo / hi are [m][d].                            */
void oracle_local_bounds_csr(const int64_t *row_ptr, const int32_t *col, const double *x, int d,
                             const int64_t *pool, uint64_t m, const double *lb, const double *ub,
                             double *lo, double *hi)
{
    for (uint64_t q = 0; q < m; ++q) {
        const int64_t v = pool[q];
        double *l = lo + q * (uint64_t)d, *h = hi + q * (uint64_t)d;
        for (int i = 0; i < d; ++i) {
            l[i] = lb[i];
            h[i] = ub[i];
        }
        for (int64_t k = row_ptr[v]; k < row_ptr[v + 1]; ++k) {
            const double *xn = x + (uint64_t)col[k] * (uint64_t)d;
            for (int i = 0; i < d; ++i) {
                const double xi = xn[i], xv = x[(uint64_t)v * (uint64_t)d + i];
                if (xi < xv && xi > l[i]) l[i] = xi;
                if (xi > xv && xi < h[i]) h[i] = xi;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* a8/a9. whole-generation simplicial complex, closed form (verified */
/* there against Complex.triangulate/refine_all, _complex.py:192-472,514-976)             */
/*   lattice resolution L = 2^(gen+1), K = 2^gen cells per axis                           */
/*   grid vertex   k in [0,K]^d   (lattice point 2k)    id = sum k_i (K+1)^i              */
/*   centroid      c in [0,K)^d   (lattice point 2c+1)  id = (K+1)^d + sum c_i K^i        */
/* ------------------------------------------------------------------------------------ */

/* per-dimension coordinate table by recursive midpointing with the reference's exact
 * expression (v2 - v1)/2.0 + v1 (_complex.py:1007); T has L+1 entries */
void oracle_lattice_table(double lb, double ub, int gen, double *T)
{
    int L = 1 << (gen + 1);
    T[0] = lb;
    T[L] = ub;
    for (int step = L; step >= 2; step >>= 1)
        for (int i = 0; i + step <= L; i += step)
            T[i + step / 2] = (T[i + step] - T[i]) / 2.0 + T[i];
}

uint64_t oracle_lattice_nvert(int d, int gen, uint64_t *ngrid_out)
{
    uint64_t K = 1ull << gen, ng = 1, nc = 1;
    for (int i = 0; i < d; ++i) { ng *= (K + 1); nc *= K; }
    if (ngrid_out) *ngrid_out = ng;
    return ng + nc;
}

/* coordinates of all vertices, row-major (nvert,d); tab is [d][L+1] */
void oracle_lattice_coords(int d, int gen, const double *tab, double *x)
{
    uint64_t K = 1ull << gen, ng;
    int L = 1 << (gen + 1);
    uint64_t nv = oracle_lattice_nvert(d, gen, &ng);
    for (uint64_t id = 0; id < nv; ++id) {
        uint64_t r = id < ng ? id : id - ng;
        for (int i = 0; i < d; ++i) {
            uint64_t base = id < ng ? K + 1 : K;
            uint64_t ki = r % base;
            r /= base;
            int p = id < ng ? (int)(2 * ki) : (int)(2 * ki + 1);
            x[id * d + i] = tab[(size_t)i * (L + 1) + p];
        }
    }
}

typedef void (*nb_visit)(uint64_t nb, void *ctx);

/* enumerate the neighbours of vertex `id` by the parity rule of Appendix B */
static void lattice_neighbours(int d, int gen, uint64_t id, nb_visit visit, void *ctx)
{
    int64_t K = 1ll << gen;
    uint64_t ng = 1;
    for (int i = 0; i < d; ++i) ng *= (uint64_t)(K + 1);
    int64_t k[64];
    if (id >= ng) { /* centroid: the 2^d corners of its cell */
        uint64_t r = id - ng;
        for (int i = 0; i < d; ++i) { k[i] = (int64_t)(r % (uint64_t)K); r /= (uint64_t)K; }
        for (uint64_t mask = 0; mask < (1ull << d); ++mask) {
            uint64_t nb = 0, mul = 1;
            for (int i = 0; i < d; ++i) {
                nb += (uint64_t)(k[i] + (int64_t)((mask >> i) & 1u)) * mul;
                mul *= (uint64_t)(K + 1);
            }
            visit(nb, ctx);
        }
        return;
    }
    uint64_t r = id;
    for (int i = 0; i < d; ++i) { k[i] = (int64_t)(r % (uint64_t)(K + 1)); r /= (uint64_t)(K + 1); }
    /* centroids of the (up to 2^d) cells around the grid vertex */
    for (uint64_t mask = 0; mask < (1ull << d); ++mask) {
        uint64_t nb = ng, mul = 1;
        int ok = 1;
        for (int i = 0; i < d; ++i) {
            int64_t c = k[i] - (int64_t)((mask >> i) & 1u);
            if (c < 0 || c >= K) { ok = 0; break; }
            nb += (uint64_t)c * mul;
            mul *= (uint64_t)K;
        }
        if (ok) visit(nb, ctx);
    }
    /* grid-grid: non-empty S within the odd-index axes P or within the even-index axes M,
     * every sign choice, |S| < d, staying inside [0,K] */
    for (int cls = 0; cls < 2; ++cls) {
        int ax[64], na = 0;
        for (int i = 0; i < d; ++i)
            if ((int)(k[i] & 1) == (cls == 0 ? 1 : 0)) ax[na++] = i;
        if (na == 0) continue;
        uint64_t combos = 1;
        for (int i = 0; i < na; ++i) combos *= 3;
        for (uint64_t t = 1; t < combos; ++t) { /* base-3 digits: 0 stay, 1 plus, 2 minus */
            uint64_t tt = t;
            int64_t kk[64];
            memcpy(kk, k, sizeof(int64_t) * (size_t)d);
            int moved = 0, ok = 1;
            for (int i = 0; i < na; ++i) {
                int dg = (int)(tt % 3);
                tt /= 3;
                if (dg == 0) continue;
                ++moved;
                kk[ax[i]] += dg == 1 ? 1 : -1;
                if (kk[ax[i]] < 0 || kk[ax[i]] > K) { ok = 0; break; }
            }
            if (!ok || moved >= d) continue;
            uint64_t nb = 0, mul = 1;
            for (int i = 0; i < d; ++i) { nb += (uint64_t)kk[i] * mul; mul *= (uint64_t)(K + 1); }
            visit(nb, ctx);
        }
    }
}

struct count_ctx { int64_t n; };
static void visit_count(uint64_t nb, void *ctx) { (void)nb; ((struct count_ctx *)ctx)->n++; }
struct fill_ctx { int32_t *col; int64_t pos; };
static void visit_fill(uint64_t nb, void *ctx)
{
    struct fill_ctx *c = (struct fill_ctx *)ctx;
    c->col[c->pos++] = (int32_t)nb;
}
static int cmp_i32(const void *a, const void *b)
{
    int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (x > y) - (x < y);
}

/* explicit CSR of the closed-form complex (small d/gen only); call with col == NULL to get
 * row_ptr / nnz first */
int64_t oracle_lattice_csr(int d, int gen, int64_t *row_ptr, int32_t *col)
{
    uint64_t nv = oracle_lattice_nvert(d, gen, NULL);
    row_ptr[0] = 0;
    for (uint64_t id = 0; id < nv; ++id) {
        struct count_ctx cc = {0};
        lattice_neighbours(d, gen, id, visit_count, &cc);
        row_ptr[id + 1] = row_ptr[id] + cc.n;
    }
    if (col) {
        for (uint64_t id = 0; id < nv; ++id) {
            struct fill_ctx fc = {col, row_ptr[id]};
            lattice_neighbours(d, gen, id, visit_fill, &fc);
            qsort(col + row_ptr[id], (size_t)(row_ptr[id + 1] - row_ptr[id]), sizeof(int32_t),
                  cmp_i32);
        }
    }
    return row_ptr[nv];
}

struct star_ctx { const double *f; double fv; int ok; };
static void visit_star(uint64_t nb, void *ctx)
{
    struct star_ctx *c = (struct star_ctx *)ctx;
    if (!(c->fv < c->f[nb])) c->ok = 0;
}

/* implicit-adjacency star test over the whole lattice complex */
void oracle_lattice_star(int d, int gen, const double *f, uint8_t *is_min)
{
    uint64_t nv = oracle_lattice_nvert(d, gen, NULL);
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t id = 0; id < (int64_t)nv; ++id) {
        struct star_ctx sc = {f, f[id], 1};
        lattice_neighbours(d, gen, (uint64_t)id, visit_star, &sc);
        is_min[id] = (uint8_t)sc.ok;
    }
}
