"""Tensor-level front end of the C ABI: the hot path on torch CUDA tensors.

PyTorch is plumbing here (device memory, streams); every function below is one or two calls into
libshgo_b200.so on ``torch.cuda.current_stream()``.  Nothing falls back to the CPU.

Reference functions replaced (paths relative to the reference checkout):
  sobol_points / eval_sobol   qmc.Sobol.random + bounds scaling + compute_sfield/feasibility_check
                              (shgo/_shgo.py:703-711,1441-1449; shgo/_shgo_lib/_vertex.py:330-352)
  csr_from_simplices          Complex.vf_to_vv (shgo/_shgo_lib/_complex.py:1053-1074)
  star_csr / lattice_star     VertexScalarField.minimiser (shgo/_shgo_lib/_vertex.py:144-151)
  compact_pool / argmin       SHGO.minimizers / find_lowest_vertex (shgo/_shgo.py:1109-1157,863-880)
"""
import numpy as np
import torch

from . import _lib, _sobol


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("shgo_b200 needs a CUDA device (B200); there is no CPU fallback")


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return t.data_ptr() if t is not None else None


def _dev(device=None):
    _require_cuda()
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


class SobolTables:
    """Device copies of the direction numbers and bounds of one problem."""

    def __init__(self, bounds, device=None):
        device = _dev(device)
        b = np.asarray(bounds, np.float64)
        self.d = b.shape[0]
        self.lb_host, self.ub_host = b[:, 0].copy(), b[:, 1].copy()
        self.sv_host = np.ascontiguousarray(_sobol.direction_numbers(self.d))
        # kernels read the tables as uint32; torch has no unsigned 32-bit arithmetic, so ship bits
        self.sv = torch.from_numpy(self.sv_host.view(np.int32).copy()).to(device)
        self.lb = torch.from_numpy(self.lb_host).to(device)
        self.ub = torch.from_numpy(self.ub_host).to(device)
        span = self.ub_host - self.lb_host
        scaled = span * 2.0 ** -30
        if not (np.all(np.isfinite(span)) and np.all(scaled * 2.0 ** 30 == span)
                and np.all((scaled == 0) | (np.abs(scaled) > 1e-290))):
            raise ValueError("bounds span not exactly scalable by 2^-30; unsupported")


MAX_FUSED_DIM = 32      # kMaxDim of the kernels (tables live in shared memory)


def soa_ld(x_soa):
    """Leading dimension (elements between consecutive coordinate rows) of a (d, n) coordinate-major
    tensor, as the C ABI wants it (`ld >= n`, include/shgo_b200.h).  torch reports an arbitrary
    stride for a dimension of size 1 (`(n, 1).t().contiguous()` keeps stride(0) == 1), so a single
    row, or an empty tensor, gets ld = max(n, 1): the value is never used to step between rows."""
    assert x_soa.dim() == 2 and (x_soa.shape[1] <= 1 or x_soa.stride(1) == 1), "rows must be unit-stride"
    d, n = x_soa.shape
    if d <= 1 or n == 0:
        return max(n, 1)
    ld = x_soa.stride(0)
    if ld < n:
        raise ValueError("coordinate rows overlap: stride(0) = %d < n = %d" % (ld, n))
    return ld


def sobol_points(tab, k0, n, layout="soa", out=None):
    """Points k0..k0+n-1 scaled to the bounds, bit-identical to the reference's `self.C`.

    layout 'soa' -> (d, n) coordinate-major; 'aos' -> (n, d) like the reference.  More than 32
    axes are generated 32 at a time (the direction numbers of an axis do not depend on the others)."""
    d = tab.d
    dev = tab.sv.device
    if d > MAX_FUSED_DIM:
        x = torch.empty((d, n), dtype=torch.float64, device=dev)
        for j0 in range(0, d, MAX_FUSED_DIM):
            j1 = min(d, j0 + MAX_FUSED_DIM)
            _lib.call("shgo_b200_sobol_fill", _ptr(tab.sv[j0:j1]), j1 - j0, k0, n, _ptr(tab.lb[j0:j1]),
                      _ptr(tab.ub[j0:j1]), _ptr(x[j0:j1]), max(n, 1), 0, _stream())
        if layout == "soa":
            return x
        return x.t().contiguous()
    if layout == "soa":
        x = out if out is not None else torch.empty((d, n), dtype=torch.float64, device=dev)
        _lib.call("shgo_b200_sobol_fill", _ptr(tab.sv), d, k0, n, _ptr(tab.lb), _ptr(tab.ub), _ptr(x),
                  soa_ld(x), 0, _stream())
    else:
        x = out if out is not None else torch.empty((n, d), dtype=torch.float64, device=dev)
        _lib.call("shgo_b200_sobol_fill", _ptr(tab.sv), d, k0, n, _ptr(tab.lb), _ptr(tab.ub), _ptr(x),
                  0, 1, _stream())
    return x


def eval_sobol(functor, gset, tab, k0, n, f=None, feasible=None, want_feasible=True):
    """Fused generate -> scale -> evaluate -> mask over Sobol indices [k0, k0+n)."""
    dev = tab.sv.device
    if f is None:
        f = torch.empty(n, dtype=torch.float64, device=dev)
    if feasible is None and want_feasible:
        feasible = torch.empty(n, dtype=torch.uint8, device=dev)
    _lib.call("shgo_b200_eval_sobol", functor, gset, tab.d, k0, n, _ptr(tab.sv), _ptr(tab.lb),
              _ptr(tab.ub), _ptr(f), _ptr(feasible), _stream())
    return f, feasible


def blockmin_layout(functor, gset):
    """log2 of the evaluation kernel's points per thread (fixes which 32 vertices form a group of the
    remote-phase filter, see include/shgo_b200.h), -1 for an unregistered pair"""
    return _lib.lib().shgo_b200_blockmin_layout(functor, gset)


def blockmin_group(i, plog):
    """group index of vertex / output index i (int or tensor) in the block_min vector"""
    tl = 8 + plog
    return ((i >> tl) << (tl - 5)) + ((i & 255) >> (5 - plog))


def eval_sobol_blockmin(functor, gset, tab, k0, n, f, feasible, block_min):
    """eval_sobol that also fills block_min (the whole n_total / 32 vector) with the group minima of
    its outputs (the filter of the sharded star test's remote phase); k0, n multiples of the tile."""
    _lib.call("shgo_b200_eval_sobol_blockmin", functor, gset, tab.d, k0, n, _ptr(tab.sv), _ptr(tab.lb),
              _ptr(tab.ub), _ptr(f), _ptr(feasible), _ptr(block_min), _stream())
    return f, feasible


def eval_points(functor, gset, x_soa, f=None, feasible=None):
    """Evaluate a registered functor on caller-supplied points, x_soa = (d, n) float64."""
    assert x_soa.dtype == torch.float64 and x_soa.dim() == 2
    d, n = x_soa.shape
    if f is None:
        f = torch.empty(n, dtype=torch.float64, device=x_soa.device)
    if feasible is None:
        feasible = torch.empty(n, dtype=torch.uint8, device=x_soa.device)
    _lib.call("shgo_b200_eval_points", functor, gset, d, n, _ptr(x_soa), soa_ld(x_soa),
              _ptr(f), _ptr(feasible), _stream())
    return f, feasible


def mask_infeasible(f, g=None, feasible=None):
    """In place: f=+inf where any g<0 or f is NaN.  g = (m, n) float64 or None."""
    n = f.numel()
    m = 0 if g is None else g.shape[0]
    if g is not None:
        g = g.contiguous()
    if feasible is None:
        feasible = torch.empty(n, dtype=torch.uint8, device=f.device)
    _lib.call("shgo_b200_mask_infeasible", _ptr(f), n, _ptr(g), m, _ptr(feasible), _stream())
    return f, feasible


def sampling_subspace(x_soa, g):
    """SHGO.sampling_subspace (_shgo.py:1452-1463) on the device: x_soa (d, n) sample points, g (m, n)
    constraint values.  -> (kept row indices int64, kept points as a C-order (m', d) tensor); a point
    survives iff every g >= 0.0 (NaN drops it)."""
    d, n = x_soa.shape
    dev = x_soa.device
    keep = torch.empty(n, dtype=torch.uint8, device=dev)
    m = 0 if g is None else g.shape[0]
    if g is not None:
        g = g.contiguous()
        assert g.dtype == torch.float64 and g.shape[1] == n
    _lib.call("shgo_b200_subspace_mask", _ptr(g), m, n, _ptr(keep), _stream())
    idx, cnt = compact_pool(keep)
    k = int(cnt.item())
    idx = idx[:k]
    out = torch.empty((k, d), dtype=torch.float64, device=dev)
    _lib.call("shgo_b200_gather_rows", _ptr(x_soa), soa_ld(x_soa), d, _ptr(idx), k, _ptr(out), _stream())
    return idx, out


def pool_rank(pool, f, x_soa=None, xl=None, by="f", xref=None):
    """The minimiser pool after SHGO.minimizers' post-processing (_shgo.py:1116-1122, 1213-1219) or
    g_topograph (_shgo.py:1227-1243): `pool` (vertex ids, the reference's cache order) without the
    vertices whose coordinates equal a row of `xl`, ordered by f (by='f') or by the euclidean distance
    from `xref` (by='distance'); equal keys keep the pool's order.  -> int64 tensor of vertex ids."""
    dev = f.device
    pool = pool.to(device=dev, dtype=torch.int64).contiguous()
    m = pool.numel()
    d = x_soa.shape[0] if x_soa is not None else 1
    k = 0
    if xl is not None and len(xl):
        xl = torch.as_tensor(np.asarray(xl, np.float64).reshape(-1, d), device=dev).contiguous()
        k = xl.shape[0]
    else:
        xl = None
    mode = {"f": 0, "distance": 1}[by]
    if xref is not None:
        xref = torch.as_tensor(np.asarray(xref, np.float64).reshape(d), device=dev).contiguous()
    ranked = torch.empty(max(m, 1), dtype=torch.int64, device=dev)
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    ws = _lib.lib().shgo_b200_pool_rank_workspace(m)
    tmp = torch.empty(ws, dtype=torch.uint8, device=dev)
    _lib.call("shgo_b200_pool_rank", _ptr(pool), m, _ptr(f), _ptr(x_soa), soa_ld(x_soa) if x_soa is not None else 0,
              d, _ptr(xl), k, mode, _ptr(xref), _ptr(ranked), _ptr(count), _ptr(tmp), ws, _stream())
    return ranked[: int(count.item())]


def csr_from_simplices(simplices, n):
    """Symmetric CSR (row_ptr int64 [n+1], col int32 [nnz]) of the reference's vertex graph."""
    assert simplices.dtype == torch.int32 and simplices.dim() == 2
    simplices = simplices.contiguous()
    S, dp1 = simplices.shape
    dev = simplices.device
    pairs = 3 * S if dp1 >= 3 else S
    row_ptr = torch.empty(n + 1, dtype=torch.int64, device=dev)
    col = torch.empty(max(2 * pairs, 4), dtype=torch.int32, device=dev)
    nnz = torch.zeros(1, dtype=torch.int64, device=dev)
    ws = _lib.lib().shgo_b200_csr_workspace(S, dp1, n)
    tmp = torch.empty(ws, dtype=torch.uint8, device=dev)
    _lib.call("shgo_b200_csr_from_simplices", _ptr(simplices), S, dp1, n, _ptr(row_ptr), _ptr(col),
              col.numel(), _ptr(nnz), _ptr(tmp), ws, _stream())
    count = int(nnz.item())
    if count < 0:          # ~0 as int64: the kernel saw a vertex index outside [0, n)
        raise ValueError("simplices refer to a vertex outside [0, %d)" % n)
    return row_ptr, col[:count]


def morton_order(x_soa, bounds):
    """Z-order numbering of the points x_soa = (d, n): (order, newid) int32 tensors with
    order[new] = old and newid[old] = new."""
    d, n = x_soa.shape
    dev = x_soa.device
    lb, ub = _bounds_vectors(bounds, d, dev)
    order = torch.empty(n, dtype=torch.int32, device=dev)
    newid = torch.empty(n, dtype=torch.int32, device=dev)
    ws = _lib.lib().shgo_b200_morton_workspace(n)
    tmp = torch.empty(ws, dtype=torch.uint8, device=dev)
    _lib.call("shgo_b200_morton_order", _ptr(x_soa), soa_ld(x_soa), d, n, _ptr(lb), _ptr(ub), _ptr(order),
              _ptr(newid), _ptr(tmp), ws, _stream())
    return order, newid


def relabel3(simplices, n, newid=None):
    """(S, 3) int32: the first three slots of every simplex (all that vf_to_vv reads), renumbered."""
    assert simplices.dtype == torch.int32 and simplices.dim() == 2
    simplices = simplices.contiguous()
    S, dp1 = simplices.shape
    out = torch.empty((S, 3), dtype=torch.int32, device=simplices.device)
    _lib.call("shgo_b200_relabel3", _ptr(simplices), S, dp1, n, _ptr(newid), _ptr(out), _stream())
    return out


def gather(src, idx, out=None):
    """out[i] = src[idx[i]] (float64 or uint8 src, int32 idx)"""
    n = idx.numel()
    if out is None:
        out = torch.empty(n, dtype=src.dtype, device=src.device)
    name = {torch.float64: "shgo_b200_gather_f64", torch.uint8: "shgo_b200_gather_u8"}[src.dtype]
    _lib.call(name, _ptr(src), _ptr(idx), n, _ptr(out), _stream())
    return out


def scatter_u8(src, idx, out):
    """out[idx[i]] = src[i]"""
    _lib.call("shgo_b200_scatter_u8", _ptr(src), _ptr(idx), idx.numel(), _ptr(out), _stream())
    return out


def csr_edges3(row_ptr, col):
    """(nnz / 2, 3) int32: every undirected edge of a symmetric CSR once, as a degenerate simplex
    [a, b, b]; together with new simplices they give csr_from_simplices the union of the meshes."""
    n = row_ptr.numel() - 1
    cap = col.numel() // 2
    out = torch.empty((max(cap, 1), 3), dtype=torch.int32, device=row_ptr.device)
    count = torch.zeros(1, dtype=torch.int64, device=row_ptr.device)
    _lib.call("shgo_b200_csr_edges3", _ptr(row_ptr), _ptr(col), n, _ptr(out), cap, _ptr(count), _stream())
    return out[:cap]


def row_present(row_ptr):
    """uint8 [n]: 1 where the row is non-empty (the vertex exists in the reference's cache)"""
    n = row_ptr.numel() - 1
    out = torch.empty(n, dtype=torch.uint8, device=row_ptr.device)
    _lib.call("shgo_b200_row_present", _ptr(row_ptr), n, _ptr(out), _stream())
    return out


def first_seen(simplices, n):
    """int64 [n]: position at which vf_to_vv first touches each vertex (-1 = never)."""
    assert simplices.dtype == torch.int32 and simplices.dim() == 2
    simplices = simplices.contiguous()
    out = torch.empty(n, dtype=torch.int64, device=simplices.device)
    _lib.call("shgo_b200_first_seen", _ptr(simplices), simplices.shape[0], simplices.shape[1], n,
              _ptr(out), _stream())
    return out


def star_csr(row_ptr, col, f, row0=0, rows=None, is_min=None, feasible=None, nfev=None):
    """Star test of rows [row0, row0+rows).  With `nfev` (1-element int64 tensor) the same pass
    also counts the rows that exist and are feasible (`feasible` indexed like is_min)."""
    n = row_ptr.numel() - 1
    rows = n - row0 if rows is None else rows
    if is_min is None:
        is_min = torch.empty(rows, dtype=torch.uint8, device=f.device)
    hint = col.numel() if rows == n else 0
    if nfev is None:
        _lib.call("shgo_b200_star_csr", _ptr(row_ptr), _ptr(col), _ptr(f), row0, rows, hint,
                  _ptr(is_min), _stream())
    else:
        _lib.call("shgo_b200_star_csr_nfev", _ptr(row_ptr), _ptr(col), _ptr(f), _ptr(feasible), row0,
                  rows, hint, _ptr(is_min), _ptr(nfev), _stream())
    return is_min


def star_csr_shard(row_ptr_local, col_local, row0, f_all, feasible, is_min, nfev, off0=None):
    """Star test of one GPU's rows against the whole (gathered) f vector.  row_ptr_local is the
    slice row_ptr[row0 : row0+rows+1] (global offsets), col_local the matching slice of col; the C
    ABI addresses the whole graph, so the slices are passed by their virtual bases."""
    rows = row_ptr_local.numel() - 1
    if off0 is None:     # pass it in to keep the call free of host synchronisation (graph capture)
        off0 = int(row_ptr_local[0].item())
    _lib.call("shgo_b200_star_csr_nfev", row_ptr_local.data_ptr() - row0 * 8,
              col_local.data_ptr() - off0 * 4, _ptr(f_all), _ptr(feasible), row0, rows,
              col_local.numel(), _ptr(is_min), _ptr(nfev), _stream())
    return is_min


def star_csr_local(row_ptr_local, col_local, own0, f_own, feasible, flags, nfev, off0=None, r0=None,
                   nrows=None, cand_ws=None, remote_at_ends=False, max_warps=0, accumulate_nfev=False):
    """Multi-GPU phase 1: rows [r0, r0+nrows) of this GPU's shard (default: all of it) against its
    OWN shard of f; flags 0 / 1 / 2 (candidate).  feasible / flags are whole-shard tensors.
    remote_at_ends: promise that other GPUs' vertices sit at the two ends of every row (faster walk)."""
    own_rows = row_ptr_local.numel() - 1
    if off0 is None:
        off0 = int(row_ptr_local[0].item())
    r0 = own0 if r0 is None else r0
    nrows = own_rows if nrows is None else nrows
    sh = r0 - own0
    hint = col_local.numel() * nrows // max(own_rows, 1)
    _lib.call("shgo_b200_star_csr_local", row_ptr_local.data_ptr() - own0 * 8,
              col_local.data_ptr() - off0 * 4, _ptr(f_own),
              feasible.data_ptr() + sh if feasible is not None else None, own0, own_rows, r0, nrows, hint,
              (1 if remote_at_ends else 0) | (2 if accumulate_nfev else 0) | (int(max_warps) << 8), flags.data_ptr() + sh, _ptr(cand_ws),
              cand_ws.numel() if cand_ws is not None else 0, _ptr(nfev),
              _stream())
    return flags


def star_csr_remote(row_ptr_local, col_local, own0, peer_table, f_own_ptr, nshards, rows_per_shard,
                    my_shard, flags, tmp, off0=None, r0=None, nrows=None, remote_at_ends=False,
                    cand_ws=None, block_min=None, layout_plog=0):
    """Multi-GPU phase 2: candidates' remote neighbours, f read in place from the peer-mapped shards.
    remote_at_ends: promise that remote neighbours form a prefix/suffix of every row (true for
    ascending rows); enables the two-ends lookup."""
    own_rows = row_ptr_local.numel() - 1
    if off0 is None:
        off0 = int(row_ptr_local[0].item())
    r0 = own0 if r0 is None else r0
    nrows = own_rows if nrows is None else nrows
    # block_min: the gathered block minima of ALL shards (float64 [n / 32]): remote reads it rules out are skipped
    _lib.call("shgo_b200_star_csr_remote_filtered", row_ptr_local.data_ptr() - own0 * 8,
              col_local.data_ptr() - off0 * 4, _ptr(peer_table), f_own_ptr, nshards, rows_per_shard,
              my_shard, r0, nrows, 1 if remote_at_ends else 0, flags.data_ptr() + (r0 - own0), _ptr(cand_ws),
              nrows, _ptr(tmp), tmp.numel() if tmp is not None else 0, _ptr(block_min), layout_plog, _stream())
    return flags


def star_csr_fused(row_ptr_local, col_local, own0, peer_table, f_own, feasible, nshards, rows_per_shard,
                   my_shard, ready, epoch, flags, cand_ws, nfev, off0=None, r0=None, nrows=None):
    """Multi-GPU phases 1 + 2 in one kernel (remote_at_ends promise required): candidates whose peers
    have published their shard (ready[q] >= epoch) read the remote f values in place over NVLink
    from inside the kernel; the rest keep flag 2 and are listed in cand_ws for star_csr_remote."""
    own_rows = row_ptr_local.numel() - 1
    if off0 is None:
        off0 = int(row_ptr_local[0].item())
    r0 = own0 if r0 is None else r0
    nrows = own_rows if nrows is None else nrows
    sh = r0 - own0
    hint = col_local.numel() * nrows // max(own_rows, 1)
    _lib.call("shgo_b200_star_csr_fused", row_ptr_local.data_ptr() - own0 * 8, col_local.data_ptr() - off0 * 4,
              _ptr(peer_table), _ptr(f_own), feasible.data_ptr() + sh if feasible is not None else None,
              nshards, rows_per_shard, my_shard, r0, nrows, hint, _ptr(ready), epoch,
              flags.data_ptr() + sh, _ptr(cand_ws), cand_ws.numel(), _ptr(nfev), _stream())
    return flags


def publish_ready(flag_table, npeers, me, epoch):
    """Store `epoch` into every peer's ready[me] (flag_table: device int64 tensor of the peers' flag
    array addresses), stream-ordered after the evaluation that wrote this GPU's shard of f."""
    _lib.call("shgo_b200_publish_ready", _ptr(flag_table), npeers, me, epoch, _stream())


class PoolWorkspace:
    """Reusable buffers for compact_pool / argmin / nfev (no allocation inside timed regions)."""

    def __init__(self, n, cap, device):
        self.cap = cap
        self.idx = torch.empty(max(cap, 1), dtype=torch.int64, device=device)
        self.count = torch.zeros(2, dtype=torch.int64, device=device)
        self.ws = _lib.lib().shgo_b200_compact_workspace(n)
        self.tmp = torch.empty(max(self.ws, 16 * 1024), dtype=torch.uint8, device=device)


def compact_pool(is_min, base=0, cap=None, work=None):
    """Ascending indices (base + i) of the set flags.  Returns (idx tensor view, count tensor)."""
    n = is_min.numel()
    if work is None:
        work = PoolWorkspace(n, n if cap is None else cap, is_min.device)
    _lib.call("shgo_b200_compact_pool", _ptr(is_min), n, base, _ptr(work.idx), work.cap,
              _ptr(work.count), _ptr(work.tmp), work.tmp.numel(), _stream())
    return work.idx, work.count[0]


def argmin(f, present=None, rank=None):
    """(index, value) of the lowest finite f among the present vertices; ties go to the lowest
    rank (int64 per vertex: the reference's cache order) when given, else to the lowest index."""
    dev = f.device
    idx = torch.empty(1, dtype=torch.int64, device=dev)
    val = torch.empty(1, dtype=torch.float64, device=dev)
    tmp = torch.empty(32 * 1024, dtype=torch.uint8, device=dev)
    _lib.call("shgo_b200_argmin_ranked", _ptr(f), _ptr(present), _ptr(rank), f.numel(), _ptr(idx), _ptr(val),
              _ptr(tmp), tmp.numel(), _stream())
    return int(idx.item()), float(val.item())


def count_nfev(row_ptr, feasible=None, out=None):
    if out is None:
        out = torch.zeros(1, dtype=torch.int64, device=row_ptr.device)
    _lib.call("shgo_b200_count_nfev", _ptr(row_ptr), _ptr(feasible), row_ptr.numel() - 1, _ptr(out), _stream())
    return out


def hypercube_csr(n, device=None):
    """Synthetic CSR of BASELINE.md section 5 (NOT a Delaunay graph): vertex i ~ i XOR 2^j,
    j < log2 n; int64 row_ptr / int32 col exactly as a qhull-derived CSR would be stored."""
    device = _dev(device)
    m = int(n).bit_length() - 1
    assert 1 << m == n, "n must be a power of two"
    row_ptr = torch.arange(0, (n + 1) * m, m, dtype=torch.int64, device=device)
    bits = (1 << torch.arange(m, dtype=torch.int32, device=device))[None, :]
    col = torch.empty((n, m), dtype=torch.int32, device=device)
    step = 1 << 24
    for s in range(0, n, step):
        e = min(n, s + step)
        i = torch.arange(s, e, dtype=torch.int32, device=device)[:, None]
        torch.bitwise_xor(i, bits, out=col[s:e])
    return row_ptr, col.reshape(-1)


# ---- whole-generation simplicial lattice ------------------------------------------------------
def lattice_eval(functor, gset, d, gen, tab, v0, nv, f=None, feasible=None):
    dev = tab.device
    if f is None:
        f = torch.empty(nv, dtype=torch.float64, device=dev)
    if feasible is None:
        feasible = torch.empty(nv, dtype=torch.uint8, device=dev)
    _lib.call("shgo_b200_lattice_eval", functor, gset, d, gen, _ptr(tab), v0, nv, _ptr(f),
              _ptr(feasible), _stream())
    return f, feasible


LATTICE_GRANULE = 4096   # default ownership granule of the multi-GPU lattice path (vertices; measured best at N = 8)


def lattice_eval_push(functor, gset, d, gen, tab, v0, nv, f, feasible, peer_ptrs, part=0, nparts=1,
                      granule=LATTICE_GRANULE):
    """lattice_eval of the granules owned by `part` of `nparts` (granule-cyclic), whose values are
    also stored into the peers' copies of f (multi-GPU): f is the rank's slice for vertices v0..,
    peer_ptrs the integer device addresses of the same position in every peer's vector (CUDA-IPC
    mappings).  Entries of granules owned by other parts are left untouched."""
    import ctypes
    arr = (ctypes.c_void_p * max(len(peer_ptrs), 1))(*peer_ptrs)
    _lib.call("shgo_b200_lattice_eval_push", functor, gset, d, gen, _ptr(tab), v0, nv, _ptr(f),
              _ptr(feasible), arr, len(peer_ptrs), part, nparts, granule, _stream())
    return f, feasible


def lattice_coords(d, gen, tab, v0, nv):
    x = torch.empty((d, nv), dtype=torch.float64, device=tab.device)
    _lib.call("shgo_b200_lattice_coords", d, gen, _ptr(tab), v0, nv, _ptr(x), max(nv, 1), _stream())
    return x


def lattice_star(d, gen, f, v0=0, nv=None, is_min=None, tmp=None, part=0, nparts=1, granule=LATTICE_GRANULE):
    """part / nparts: classify only the granules owned by `part` (the other flags come back 0)"""
    nv = f.numel() - v0 if nv is None else nv
    if is_min is None:
        is_min = torch.empty(nv, dtype=torch.uint8, device=f.device)
    ws = _lib.lib().shgo_b200_lattice_star_workspace(nv)
    if tmp is None or tmp.numel() < ws:
        tmp = torch.empty(ws, dtype=torch.uint8, device=f.device)
    _lib.call("shgo_b200_lattice_star_part", d, gen, _ptr(f), v0, nv, part, nparts, granule, _ptr(is_min), _ptr(tmp),
              tmp.numel(), _stream())
    return is_min


# ---- local bounds of pool vertices ------------------------------------
def _bounds_vectors(bounds, d, dev):
    b = torch.as_tensor(np.asarray(bounds, np.float64).reshape(d, 2), device=dev)
    return b[:, 0].contiguous(), b[:, 1].contiguous()


def local_bounds_csr(row_ptr, col, x_soa, pool, bounds):
    """construct_lcb_simplicial (_shgo.py:1246-1270) of the vertices `pool` of an explicit graph;
    x_soa is (d, ld) coordinate-major.  -> lo, hi of shape (m, d)."""
    d = x_soa.shape[0]
    ld = soa_ld(x_soa)
    dev = x_soa.device
    pool = pool.to(device=dev, dtype=torch.int64).contiguous()
    m = pool.numel()
    lb, ub = _bounds_vectors(bounds, d, dev)
    lo = torch.empty((m, d), dtype=torch.float64, device=dev)
    hi = torch.empty((m, d), dtype=torch.float64, device=dev)
    _lib.call("shgo_b200_local_bounds_csr", _ptr(row_ptr), _ptr(col), _ptr(x_soa), ld, d, _ptr(pool), m,
              _ptr(lb), _ptr(ub), _ptr(lo), _ptr(hi), _stream())
    return lo, hi


def lattice_local_bounds(d, gen, tab, pool, bounds):
    dev = tab.device
    pool = pool.to(device=dev, dtype=torch.int64).contiguous()
    m = pool.numel()
    lb, ub = _bounds_vectors(bounds, d, dev)
    lo = torch.empty((m, d), dtype=torch.float64, device=dev)
    hi = torch.empty((m, d), dtype=torch.float64, device=dev)
    _lib.call("shgo_b200_lattice_local_bounds", d, gen, _ptr(tab), _ptr(pool), m, _ptr(lb), _ptr(ub),
              _ptr(lo), _ptr(hi), _stream())
    return lo, hi


def fp64_peak_tflops(iters=4096, repeats=5):
    """Measured FP64 DFMA throughput of the current device (TFLOP/s)."""
    _require_cuda()
    import ctypes
    sink = torch.zeros(1, dtype=torch.float64, device="cuda")
    flops = ctypes.c_uint64(0)
    best = 0.0
    for _ in range(repeats):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call("shgo_b200_fp64_peak", iters, _ptr(sink), ctypes.byref(flops), _stream())
        e1.record()
        e1.synchronize()
        best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best
