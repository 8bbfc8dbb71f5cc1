"""TEST INFRASTRUCTURE: a CPU stand-in for shgo_b200.hotpath built on the oracle.

Lets the `-m "not gpu"` suite exercise the HOST logic of DeviceComplex / the shgo() wrapper (lazy
vertex materialisation, star() bookkeeping, mode switching, mesh unions, host-callable route) in
a container without a GPU.  It is never imported by the product; the GPU suite runs the same tests
against the real kernels.
"""
import numpy as np
import torch

from oracle import oracle as orc

_FN = {v: k for k, v in orc.FUNCTOR_IDS.items()}
_GN = {0: None, 1: "cattle_feed", 2: "sum_le_6"}
CPU = torch.device("cpu")


class SobolTables:
    def __init__(self, bounds, device=None):
        b = np.asarray(bounds, np.float64)
        self.d = b.shape[0]
        self.lb_host, self.ub_host = b[:, 0].copy(), b[:, 1].copy()
        self.sv_host = orc.sobol_dirnums(self.d)
        self.sv = torch.zeros(1)
        self.lb, self.ub = torch.from_numpy(self.lb_host), torch.from_numpy(self.ub_host)


def sobol_points(tab, k0, n, layout="soa", out=None):
    x = orc.sobol_points(tab.sv_host, k0, n, tab.lb_host, tab.ub_host)
    return torch.from_numpy(x.T.copy() if layout == "soa" else x)


def eval_sobol(functor, gset, tab, k0, n, f=None, feasible=None, want_feasible=True):
    fo, fe = orc.sobol_eval(tab.sv_host, _FN[functor], _GN[gset], k0, n, tab.lb_host, tab.ub_host)
    fo, fe = torch.from_numpy(fo), torch.from_numpy(fe)
    if f is not None:
        f.copy_(fo)
        fo = f
    if feasible is not None:
        feasible.copy_(fe)
        fe = feasible
    return fo, fe


def eval_points(functor, gset, x_soa, f=None, feasible=None):
    fo, fe = orc.eval_points(_FN[functor], _GN[gset], np.ascontiguousarray(x_soa.numpy().T))
    return torch.from_numpy(fo), torch.from_numpy(fe)


def mask_infeasible(f, g=None, feasible=None):
    fh = f.numpy().copy()
    ok = np.ones(len(fh), bool)
    if g is not None:
        ok = ~np.any(g.numpy() < 0.0, axis=0)
    fh[~ok | np.isnan(fh)] = np.inf
    return torch.from_numpy(fh), torch.from_numpy(ok.astype(np.uint8))


def csr_from_simplices(simplices, n):
    rp, col = orc.vf_to_vv(simplices.numpy().astype(np.int32), n)
    return torch.from_numpy(rp), torch.from_numpy(col)


def morton_order(x_soa, bounds):
    """any locality-ish permutation serves the host logic: stable sort by a coarse Z-order key"""
    key = orc.morton_keys(np.ascontiguousarray(x_soa.numpy().T), np.asarray(bounds, float))
    order = np.argsort(key, kind="stable").astype(np.int32)
    newid = np.empty_like(order)
    newid[order] = np.arange(len(order), dtype=np.int32)
    return torch.from_numpy(order), torch.from_numpy(newid)


def relabel3(simplices, n, newid=None):
    s = simplices.numpy()
    s3 = s[:, :3] if s.shape[1] >= 3 else np.concatenate([s, s[:, 1:2]], 1)
    if newid is not None:
        s3 = newid.numpy()[s3]
    return torch.from_numpy(np.ascontiguousarray(s3.astype(np.int32)))


def gather(src, idx, out=None):
    return src[idx.to(torch.int64)].contiguous()


def scatter_u8(src, idx, out):
    out[idx.to(torch.int64)] = src
    return out


def sampling_subspace(x_soa, g):
    C = np.ascontiguousarray(x_soa.numpy().T)
    keep = np.ones(C.shape[0], bool) if g is None else np.all(g.numpy() >= 0.0, axis=0)
    return torch.from_numpy(np.flatnonzero(keep)), torch.from_numpy(np.ascontiguousarray(C[keep]))


def pool_rank(pool, f, x_soa=None, xl=None, by="f", xref=None):
    x = None if x_soa is None else np.ascontiguousarray(x_soa.numpy().T)
    out = orc.pool_rank(pool.numpy(), f.numpy(), x, [] if xl is None else xl, by=by, xref=xref)
    return torch.from_numpy(np.ascontiguousarray(out))


def csr_edges3(row_ptr, col):
    rp, c = row_ptr.numpy(), col.numpy()
    rows = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
    keep = rows < c
    return torch.from_numpy(np.stack([rows[keep], c[keep], c[keep]], 1).astype(np.int32))


def row_present(row_ptr):
    rp = row_ptr.numpy()
    return torch.from_numpy((rp[1:] > rp[:-1]).astype(np.uint8))


def count_nfev(row_ptr, feasible=None, out=None):
    c = orc.nfev(row_ptr.numpy(), None if feasible is None else feasible.numpy())
    return torch.tensor([int(c)], dtype=torch.int64)


def first_seen(simplices, n):
    s = simplices.numpy()
    k = min(s.shape[1], 3)
    flat = s[:, :k].reshape(-1).astype(np.int64)
    pos = (np.arange(s.shape[0])[:, None] * 3 + np.arange(k)[None, :]).reshape(-1)
    out = np.full(n, -1, np.int64)
    u, first = np.unique(flat, return_index=True)
    out[u] = pos[first]
    return torch.from_numpy(out)


def star_csr(row_ptr, col, f, row0=0, rows=None, is_min=None, feasible=None, nfev=None):
    rp, c = row_ptr.numpy(), np.ascontiguousarray(col.numpy())
    out = orc.star_csr(rp, c, f.numpy(), row0, rows)
    if nfev is not None:
        nfev[0] = orc.nfev(rp, None if feasible is None else feasible.numpy())
    return torch.from_numpy(out)


def star_csr_shard(row_ptr_local, col_local, row0, f_all, feasible, is_min, nfev, off0=None):
    rp = row_ptr_local.numpy()
    rows = len(rp) - 1
    rp0 = rp - rp[0]
    out = np.empty(rows, np.uint8)
    fh, c = f_all.numpy(), col_local.numpy()
    for r in range(rows):
        nb = c[rp0[r]:rp0[r + 1]]
        out[r] = 1 if len(nb) and np.all(fh[row0 + r] < fh[nb]) else 0
    is_min.copy_(torch.from_numpy(out))
    nfev[0] = int(np.sum((np.diff(rp) > 0) & (feasible.numpy() != 0)))
    return is_min


class PoolWorkspace:
    def __init__(self, n, cap, device):
        self.cap = cap
        self.idx = torch.empty(max(cap, 1), dtype=torch.int64)
        self.count = torch.zeros(2, dtype=torch.int64)
        self.tmp = torch.empty(16 * 1024, dtype=torch.uint8)


def compact_pool(is_min, base=0, cap=None, work=None):
    if work is not None:
        idx = orc.compact_pool(np.ascontiguousarray(is_min.numpy())) + base
        work.count[0] = len(idx)
        m = min(len(idx), work.cap)
        work.idx[:m] = torch.from_numpy(idx[:m])
        return work.idx, work.count[0]
    return _compact_pool(is_min, base)


def _compact_pool(is_min, base=0):
    idx = orc.compact_pool(np.ascontiguousarray(is_min.numpy())) + base
    return torch.from_numpy(idx), torch.tensor(len(idx))


def argmin(f, present=None, rank=None):
    fh = f.numpy()
    i = orc.argmin(fh, None if present is None else np.ascontiguousarray(present.numpy()))
    if i >= 0 and rank is not None:       # ties: the lowest rank instead of the lowest index
        tie = fh == fh[i]
        if present is not None:
            tie &= present.numpy() != 0
        cand = np.flatnonzero(tie)
        i = int(cand[np.argmin(rank.numpy()[cand])])
    return i, (float(fh[i]) if i >= 0 else float("inf"))


def _bounds_from_tab(tab):
    t = tab.numpy()
    return [(row[0], row[-1]) for row in t]


def lattice_eval(functor, gset, d, gen, tab, v0, nv, f=None, feasible=None):
    x = orc.lattice_coords(_bounds_from_tab(tab), gen)[v0:v0 + nv]
    fo, fe = orc.eval_points(_FN[functor], _GN[gset], x)
    fo, fe = torch.from_numpy(fo), torch.from_numpy(fe)
    if f is not None:
        f.copy_(fo)
        fo = f
    if feasible is not None:
        feasible.copy_(fe)
        fe = feasible
    return fo, fe


LATTICE_GRANULE = 4096


def _owned(v0, nv, part, nparts, granule):
    return ((np.arange(v0, v0 + nv) // granule) % nparts) == part


def lattice_eval_push(functor, gset, d, gen, tab, v0, nv, f, feasible, peer_ptrs, part=0, nparts=1,
                      granule=LATTICE_GRANULE):
    assert not peer_ptrs, "no peer memory on the CPU stand-in"
    fo, fe = lattice_eval(functor, gset, d, gen, tab, v0, nv)
    own = torch.from_numpy(_owned(v0, nv, part, nparts, granule))
    f[own] = fo[own]
    feasible[own] = fe[own]
    return f, feasible


def lattice_coords(d, gen, tab, v0, nv):
    x = orc.lattice_coords(_bounds_from_tab(tab), gen)[v0:v0 + nv]
    return torch.from_numpy(np.ascontiguousarray(x.T))


def lattice_star(d, gen, f, v0=0, nv=None, is_min=None, tmp=None, part=0, nparts=1, granule=LATTICE_GRANULE):
    out = orc.lattice_star(d, gen, f.numpy())
    nv = len(out) - v0 if nv is None else nv
    res = torch.from_numpy(out[v0:v0 + nv].copy())
    res[torch.from_numpy(~_owned(v0, nv, part, nparts, granule))] = 0
    if is_min is not None:
        is_min.copy_(res)
        return is_min
    return res


def local_bounds_csr(row_ptr, col, x_soa, pool, bounds):
    x = np.ascontiguousarray(x_soa.numpy().T)
    lo, hi = orc.local_bounds_csr(row_ptr.numpy(), col.numpy(), x, pool.numpy(), bounds)
    return torch.from_numpy(lo), torch.from_numpy(hi)


def lattice_local_bounds(d, gen, tab, pool, bounds):
    """independent of the product's closed form: walks the oracle's explicit lattice adjacency"""
    x = orc.lattice_coords(_bounds_from_tab(tab), gen)
    rp, col = orc.lattice_csr(d, gen)
    lo, hi = orc.local_bounds_csr(rp, col, x, pool.numpy(), bounds)
    return torch.from_numpy(lo), torch.from_numpy(hi)


# what the C ABI silently assumes of every tensor it is handed: contiguous memory of the exact type
_DTYPES = {"f": torch.float64, "f_all": torch.float64, "f_own": torch.float64, "g": torch.float64,
           "x_soa": torch.float64, "tab": torch.float64, "feasible": torch.uint8, "is_min": torch.uint8,
           "flags": torch.uint8, "present": torch.uint8, "row_ptr": torch.int64, "row_ptr_local": torch.int64,
           "col": torch.int32, "col_local": torch.int32, "simplices": torch.int32, "pool": None, "nfev": torch.int64}


def _dry_run(real, args, kwargs):
    """Run the PRODUCT's marshalling code (shgo_b200.hotpath.<fn>) with `_lib.call` replaced by the
    argument contract of the C ABI (tests/abi_contract.py): what the CUDA library would reject with
    SHGO_B200_ERR_ARG on a B200 fails here.  Output buffers are cloned so nothing is computed twice."""
    from shgo_b200 import _lib
    import abi_contract
    saved = _lib.call
    _lib.call = abi_contract.check
    try:
        real(*args, **kwargs)
    finally:
        _lib.call = saved


def _strict(fn, real=None):
    import functools
    import inspect
    sig = inspect.signature(fn)

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        if real is not None:
            _dry_run(real, args, kwargs)
        bound = sig.bind(*args, **kwargs)
        for name, val in bound.arguments.items():
            if isinstance(val, torch.Tensor):
                assert val.is_contiguous(), "%s(%s): non-contiguous tensor handed to the hot path" % (fn.__name__, name)
                want = _DTYPES.get(name)
                if want is not None:
                    assert val.dtype == want, "%s(%s): dtype %s, the C ABI reads %s" % (fn.__name__, name, val.dtype, want)
        return fn(*args, **kwargs)
    wrapper.__signature__ = sig
    return wrapper


def install(monkeypatch):
    """Patch shgo_b200.hotpath in place (functions are looked up as attributes at call time)."""
    from shgo_b200 import hotpath as hp
    for name in ("SobolTables", "sobol_points", "eval_sobol", "eval_points", "mask_infeasible",
                 "csr_from_simplices", "csr_edges3", "row_present", "sampling_subspace", "pool_rank", "count_nfev", "morton_order", "relabel3", "gather", "scatter_u8", "first_seen", "star_csr", "star_csr_shard", "PoolWorkspace", "compact_pool", "argmin", "lattice_eval",
                 "lattice_coords", "lattice_star", "local_bounds_csr", "lattice_local_bounds", "lattice_eval_push"):
        obj = globals()[name]
        real = getattr(hp, name)
        real = getattr(real, "__wrapped_real__", real)     # installing twice keeps the product's function
        if not isinstance(obj, type):
            obj = _strict(obj, real)
            obj.__wrapped_real__ = real
        monkeypatch.setattr(hp, name, obj)
    monkeypatch.setattr(hp, "_require_cuda", lambda: None)
    monkeypatch.setattr(hp, "_dev", lambda device=None: CPU)
    monkeypatch.setattr(hp, "_stream", lambda: 0)
