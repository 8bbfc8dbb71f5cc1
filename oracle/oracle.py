"""ctypes front-end of the CPU oracle (oracle/shgo_oracle.c) plus numpy restatements.

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product never imports this module.

The numpy definitions of the objectives below are the *reference-side callables*: what a
user of the reference would hand to ``shgo(func, ...)`` (VertexCacheField.compute_sfield calls
them one vertex at a time, /root/reference/shgo/_shgo_lib/_vertex.py:338-352).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

FUNCTOR_IDS = {
    "rosenbrock": 0,
    "rastrigin": 1,
    "ackley": 2,
    "styblinski_tang": 3,
    "eggholder": 4,
    "cattle_feed": 5,
    "sphere": 6,
}
GSET_IDS = {None: 0, "none": 0, "cattle_feed": 1, "sum_le_6": 2}


def build(force=False):
    """Compile the C restatement (gcc only; no GPU, no reference checkout needed)."""
    src = os.path.join(_HERE, "shgo_oracle.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL,
                          stderr=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        c = ctypes
        vp = c.c_void_p
        L.oracle_num_threads.restype = c.c_int
        L.oracle_set_num_threads.argtypes = [c.c_int]
        L.oracle_sobol_dirnums.argtypes = [c.c_int, vp, vp, vp, vp]
        L.oracle_sobol_points.argtypes = [vp, c.c_int, c.c_uint64, c.c_uint64, vp, vp, vp]
        L.oracle_eval.argtypes = [c.c_int, c.c_int, c.c_int, c.c_uint64, vp, vp, vp]
        L.oracle_eval.restype = c.c_int
        L.oracle_sobol_eval.argtypes = [vp, c.c_int, c.c_int, c.c_int, c.c_uint64, c.c_uint64,
                                        vp, vp, vp, vp]
        L.oracle_sobol_eval.restype = c.c_int
        L.oracle_vf_to_vv.argtypes = [vp, c.c_uint64, c.c_int, c.c_uint64, vp, vp]
        L.oracle_vf_to_vv.restype = c.c_int64
        L.oracle_star_csr.argtypes = [vp, vp, vp, c.c_uint64, c.c_uint64, vp]
        L.oracle_compact_pool.argtypes = [vp, c.c_uint64, vp]
        L.oracle_compact_pool.restype = c.c_uint64
        L.oracle_argmin.argtypes = [vp, c.c_uint64, vp]
        L.oracle_argmin.restype = c.c_int64
        L.oracle_nfev.argtypes = [vp, vp, c.c_uint64]
        L.oracle_nfev.restype = c.c_uint64
        L.oracle_local_bounds_csr.argtypes = [vp, vp, vp, c.c_int, vp, c.c_uint64, vp, vp, vp, vp]
        L.oracle_lattice_table.argtypes = [c.c_double, c.c_double, c.c_int, vp]
        L.oracle_lattice_nvert.argtypes = [c.c_int, c.c_int, vp]
        L.oracle_lattice_nvert.restype = c.c_uint64
        L.oracle_lattice_coords.argtypes = [c.c_int, c.c_int, vp, vp]
        L.oracle_lattice_csr.argtypes = [c.c_int, c.c_int, vp, vp]
        L.oracle_lattice_csr.restype = c.c_int64
        L.oracle_lattice_star.argtypes = [c.c_int, c.c_int, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def num_threads():
    return int(lib().oracle_num_threads())


def set_threads(n):
    """Use n OpenMP threads regardless of OMP_NUM_THREADS (torchrun sets it to 1)."""
    lib().oracle_set_num_threads(int(n))


# ---------------------------------------------------------------------------------------
# Sobol
# ---------------------------------------------------------------------------------------
def load_joe_kuo(path, d):
    """Parse `d s a m_i` rows (the format of the reference's sobol_vec.gz)."""
    s = np.zeros(d, np.int32)
    a = np.zeros(d, np.int32)
    m = np.zeros((d, 18), np.int32)
    with open(path) as fh:
        for line in fh:
            if line.startswith("#") or not line.strip():
                continue
            t = line.split()
            j = int(t[0]) - 1
            if j >= d:
                break
            s[j], a[j] = int(t[1]), int(t[2])
            m[j, :s[j]] = [int(v) for v in t[3:3 + s[j]]]
    return s, a, m


def sobol_dirnums(d, table_path=None):
    if table_path is None:
        table_path = os.path.join(_HERE, "..", "shgo_b200", "data", "joe_kuo_dirnums.txt")
    s, a, m = load_joe_kuo(table_path, d)
    sv = np.zeros((d, 30), np.uint32)
    lib().oracle_sobol_dirnums(d, _p(s), _p(a), _p(m), _p(sv))
    return sv


def sobol_points(sv, k0, n, lb, ub):
    d = sv.shape[0]
    lb = np.ascontiguousarray(lb, np.float64)
    ub = np.ascontiguousarray(ub, np.float64)
    out = np.empty((n, d), np.float64)
    lib().oracle_sobol_points(_p(sv), d, k0, n, _p(lb), _p(ub), _p(out))
    return out


# ---------------------------------------------------------------------------------------
# objectives: the numpy callables a reference user passes as `func`
# ---------------------------------------------------------------------------------------
def np_rosenbrock(x):
    x = np.asarray(x)
    return np.sum(100.0 * (x[1:] - x[:-1] ** 2.0) ** 2.0 + (1 - x[:-1]) ** 2.0, axis=0)


def np_rastrigin(x):
    return 10 * len(x) + np.sum(x ** 2 - 10 * np.cos(2 * np.pi * x))


def np_ackley(x):
    d = len(x)
    return (-20 * np.exp(-0.2 * np.sqrt(np.sum(x ** 2) / d))
            - np.exp(np.sum(np.cos(2 * np.pi * x)) / d) + 20 + np.e)


def np_styblinski_tang(x):
    return 0.5 * np.sum(x ** 4 - 16 * x ** 2 + 5 * x)


def np_eggholder(x):
    return (-(x[1] + 47.0) * np.sin(np.sqrt(abs(x[0] / 2.0 + (x[1] + 47.0))))
            - x[0] * np.sin(np.sqrt(abs(x[0] - (x[1] + 47.0)))))


def np_cattle_feed(x):
    return 24.55 * x[0] + 26.75 * x[1] + 39 * x[2] + 40.50 * x[3]


def np_cattle_g1(x):
    return 2.3 * x[0] + 5.6 * x[1] + 11.1 * x[2] + 1.3 * x[3] - 5


def np_cattle_g2(x):
    return (12 * x[0] + 11.9 * x[1] + 41.8 * x[2] + 52.1 * x[3] - 21
            - 1.645 * np.sqrt(0.28 * x[0] ** 2 + 0.19 * x[1] ** 2
                              + 20.5 * x[2] ** 2 + 0.62 * x[3] ** 2))


def np_sphere(x):
    r = x[0] ** 2
    for i in range(1, len(x)):
        r = r + x[i] ** 2
    return r


def np_sum_le_6(x):
    return -(np.sum(x, axis=0) - 6.0)


NP_OBJECTIVES = {
    "rosenbrock": np_rosenbrock,
    "rastrigin": np_rastrigin,
    "ackley": np_ackley,
    "styblinski_tang": np_styblinski_tang,
    "eggholder": np_eggholder,
    "cattle_feed": np_cattle_feed,
    "sphere": np_sphere,
}
NP_CONSTRAINTS = {
    None: (),
    "none": (),
    "cattle_feed": (np_cattle_g1, np_cattle_g2),
    "sum_le_6": (np_sum_le_6,),
}


def eval_points(functor, gset, x):
    """C restatement of compute_sfield/feasibility_check over (n,d) points."""
    x = np.ascontiguousarray(x, np.float64)
    n, d = x.shape
    f = np.empty(n, np.float64)
    feas = np.empty(n, np.uint8)
    rc = lib().oracle_eval(FUNCTOR_IDS[functor], GSET_IDS[gset], d, n, _p(x), _p(f), _p(feas))
    assert rc == 0
    return f, feas


def sobol_eval(sv, functor, gset, k0, n, lb, ub):
    d = sv.shape[0]
    lb = np.ascontiguousarray(lb, np.float64)
    ub = np.ascontiguousarray(ub, np.float64)
    f = np.empty(n, np.float64)
    feas = np.empty(n, np.uint8)
    rc = lib().oracle_sobol_eval(_p(sv), FUNCTOR_IDS[functor], GSET_IDS[gset], d, k0, n,
                                 _p(lb), _p(ub), _p(f), _p(feas))
    assert rc == 0
    return f, feas


def eval_points_numpy(functor, gset, x):
    """The reference's own per-vertex evaluation order, through the numpy callables."""
    fn = NP_OBJECTIVES[functor]
    gs = NP_CONSTRAINTS[gset]
    n = x.shape[0]
    f = np.empty(n)
    feas = np.ones(n, np.uint8)
    for i in range(n):
        xi = x[i]
        ok = True
        for g in gs:
            if g(xi) < 0.0:
                ok = False
                break
        if ok:
            v = fn(xi)
            f[i] = np.inf if np.isnan(v) else v
        else:
            f[i] = np.inf
            feas[i] = 0
    return f, feas


# ---------------------------------------------------------------------------------------
# graph + classification
# ---------------------------------------------------------------------------------------
def vf_to_vv(simplices, n):
    simplices = np.ascontiguousarray(simplices, np.int32)
    S, dp1 = simplices.shape
    row_ptr = np.zeros(n + 1, np.int64)
    col = np.empty(max(1, 6 * S), np.int32)
    nnz = lib().oracle_vf_to_vv(_p(simplices), S, dp1, n, _p(row_ptr), _p(col))
    return row_ptr, col[:nnz].copy()


def star_csr(row_ptr, col, f, row0=0, rows=None):
    n = len(row_ptr) - 1
    rows = n - row0 if rows is None else rows
    out = np.empty(rows, np.uint8)
    f = np.ascontiguousarray(f, np.float64)
    lib().oracle_star_csr(_p(row_ptr), _p(col), _p(f), row0, rows, _p(out))
    return out


def compact_pool(is_min):
    idx = np.empty(len(is_min), np.int64)
    c = lib().oracle_compact_pool(_p(is_min), len(is_min), _p(idx))
    return idx[:c].copy()


def argmin(f, present=None):
    f = np.ascontiguousarray(f, np.float64)
    return int(lib().oracle_argmin(_p(f), len(f), _p(present)))


def nfev(row_ptr, feas):
    return int(lib().oracle_nfev(_p(row_ptr), _p(feas), len(row_ptr) - 1))


def local_bounds_csr(row_ptr, col, x, pool, bounds):
    """construct_lcb_simplicial of the vertices `pool`; x row-major (n,d).  -> lo, hi (m,d)."""
    x = np.ascontiguousarray(x, np.float64)
    d = x.shape[1]
    pool = np.ascontiguousarray(pool, np.int64)
    col = np.ascontiguousarray(col, np.int32)
    b = np.asarray(bounds, np.float64).reshape(d, 2)
    lb, ub = np.ascontiguousarray(b[:, 0]), np.ascontiguousarray(b[:, 1])
    lo = np.empty((len(pool), d))
    hi = np.empty((len(pool), d))
    lib().oracle_local_bounds_csr(_p(row_ptr), _p(col), _p(x), d, _p(pool), len(pool), _p(lb), _p(ub),
                                  _p(lo), _p(hi))
    return lo, hi


def csr_from_edges(edges, n):
    """Symmetric CSR (ascending rows) of an undirected edge list."""
    e = np.asarray(edges, np.int64).reshape(-1, 2)
    both = np.concatenate([e, e[:, ::-1]])
    both = both[np.lexsort((both[:, 1], both[:, 0]))]
    row_ptr = np.zeros(n + 1, np.int64)
    np.add.at(row_ptr, both[:, 0] + 1, 1)
    return np.cumsum(row_ptr), both[:, 1].astype(np.int32)


def hypercube_csr(n):
    """Synthetic CSR of BASELINE.md section 5: vertex i adjacent to i XOR 2^j, j < log2 n."""
    m = int(n).bit_length() - 1
    assert 1 << m == n
    row_ptr = (np.arange(n + 1, dtype=np.int64) * m)
    i = np.arange(n, dtype=np.int64)[:, None]
    col = (i ^ (1 << np.arange(m, dtype=np.int64))[None, :]).astype(np.int32)
    return row_ptr, col.reshape(-1)


# ---------------------------------------------------------------------------------------
# simplicial lattice (whole generations)
# ---------------------------------------------------------------------------------------
def lattice_tables(bounds, gen):
    L = 1 << (gen + 1)
    tab = np.empty((len(bounds), L + 1), np.float64)
    for i, (lo, hi) in enumerate(bounds):
        lib().oracle_lattice_table(float(lo), float(hi), gen, _p(tab[i]))
    return tab


def lattice_nvert(d, gen):
    ng = ctypes.c_uint64(0)
    nv = lib().oracle_lattice_nvert(d, gen, ctypes.byref(ng))
    return int(nv), int(ng.value)


def lattice_coords(bounds, gen):
    d = len(bounds)
    tab = lattice_tables(bounds, gen)
    nv, _ = lattice_nvert(d, gen)
    x = np.empty((nv, d), np.float64)
    lib().oracle_lattice_coords(d, gen, _p(tab), _p(x))
    return x


def lattice_csr(d, gen):
    nv, _ = lattice_nvert(d, gen)
    row_ptr = np.zeros(nv + 1, np.int64)
    nnz = lib().oracle_lattice_csr(d, gen, _p(row_ptr), None)
    col = np.empty(max(1, nnz), np.int32)
    lib().oracle_lattice_csr(d, gen, _p(row_ptr), _p(col))
    return row_ptr, col[:nnz]


def lattice_star(d, gen, f):
    nv, _ = lattice_nvert(d, gen)
    f = np.ascontiguousarray(f, np.float64)
    assert len(f) == nv
    out = np.empty(nv, np.uint8)
    lib().oracle_lattice_star(d, gen, _p(f), _p(out))
    return out


def morton_keys(x, bounds):
    """32-bit Z-order key of every point (the numbering shgo_b200_morton_order sorts by; the
    reference has no counterpart -- its vertices live in a dict, _vertex.py:304-317).  x: (n, d)."""
    x = np.asarray(x, np.float64)
    n, d = x.shape
    b = np.asarray(bounds, np.float64).reshape(d, 2)
    dk = min(d, 32)
    bits = 30 if dk == 1 else 32 // dk
    key = np.zeros(n, np.uint64)
    for j in range(dk):
        w = b[j, 1] - b[j, 0]
        t = (x[:, j] - b[j, 0]) / w if w > 0 else np.zeros(n)
        t = np.where(t == t, np.clip(t, 0.0, 1.0), 0.0)
        q = np.minimum((t * float(1 << bits)).astype(np.uint64), np.uint64((1 << bits) - 1))
        for bit in range(bits):
            key |= ((q >> np.uint64(bit)) & np.uint64(1)) << np.uint64(bit * dk + (dk - 1 - j))
    return key.astype(np.uint32)


def sampling_subspace(C, g_funcs):
    """restatement of SHGO.sampling_subspace, reference _shgo.py:1452-1463 (C: (n, d) points)"""
    idx = np.arange(C.shape[0])
    for g in g_funcs:
        keep = g(C.T) >= 0.0
        C, idx = C[keep], idx[keep]
    return idx, C


def pool_rank(pool, f, x=None, xl=(), by="f", xref=None):
    """restatement of the pool post-processing of SHGO.minimizers / sort_min_pool / g_topograph
    (reference _shgo.py:1116-1122, 1213-1219, 1227-1243) with a STABLE argsort"""
    pool = np.asarray(pool, np.int64)
    keep = np.ones(len(pool), bool)
    for xm in xl:
        keep &= ~np.all(x[pool] == np.asarray(xm), axis=1)
    pool = pool[keep]
    if by == "f":
        key = f[pool]
    else:
        from scipy import spatial
        key = spatial.distance.cdist(np.array([xref]), x[pool], "euclidean")[0]
    return pool[np.argsort(key, kind="stable")]
