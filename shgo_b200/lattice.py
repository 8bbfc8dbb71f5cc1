"""Host side of the whole-generation simplicial ("hypercube / Freudenthal") complex.

The reference grows this complex vertex by vertex (Complex.triangulate / refine_all,
shgo/_shgo_lib/_complex.py:192-472, 514-976).  When whole generations are built it has a closed
form (verified there and again in tests/ against the live reference):

  lattice resolution L = 2^(gen+1), K = 2^gen cells per axis
  grid vertex   k in [0,K]^d  -> lattice point 2k,    id = sum k_i (K+1)^i
  centroid      c in [0,K)^d  -> lattice point 2c+1,  id = (K+1)^d + sum c_i K^i
  coordinate of lattice index p on axis i = T_i[p], T_i built by recursive midpointing with the
  reference's exact expression (v2 - v1)/2.0 + v1 (split_edge, _complex.py:1007).

Only the O(d * L) coordinate tables are built here; vertices, coordinates, f and the star test
live on the device (csrc/sources.cuh LatticeSrc, csrc/lattice.cu).
"""
import numpy as np


class LatticeArtefactError(ValueError):
    """Bounds for which the REFERENCE itself builds near-duplicate vertices (the fp64 midpoint of
    an edge depends on the direction it is split from, SURVEY.md section 7 hard part 6): the closed
    form does not apply; DeviceComplex then runs the reference's own generator for the graph."""


def coordinate_tables(bounds, gen, check=True):
    """[d][2^(gen+1)+1] float64 per-axis coordinate tables."""
    L = 1 << (gen + 1)
    tab = np.empty((len(bounds), L + 1), np.float64)
    for i, (lo, hi) in enumerate(bounds):
        T = tab[i]
        T[0], T[L] = np.float64(lo), np.float64(hi)
        step = L
        while step >= 2:
            a = T[0:L:step]
            b = T[step:L + 1:step]
            mid = (b - a) / 2.0 + a
            if check and not np.array_equal(mid, (a - b) / 2.0 + b):
                raise LatticeArtefactError(
                    "bounds (%r, %r) on axis %d: midpoint depends on split direction in fp64; the "
                    "reference builds duplicate vertices here" % (lo, hi, i))
            T[step // 2:L:step] = mid
            step //= 2
    return tab


def counts(d, gen):
    """(number of vertices, number of grid vertices)"""
    K = 1 << gen
    ngrid = (K + 1) ** d
    return ngrid + K ** d, ngrid


def vertex_ids(points_idx, d, gen, grid):
    """ids from integer digits (n,d): grid digits k or centroid digits c."""
    K = 1 << gen
    base = K + 1 if grid else K
    mul = base ** np.arange(d, dtype=np.int64)
    ids = (np.asarray(points_idx, np.int64) * mul).sum(axis=-1)
    return ids if grid else ids + (K + 1) ** d


def decode(ids, d, gen):
    """lattice indices p (n,d) of vertex ids (host mirror of LatticeSrc)."""
    ids = np.asarray(ids, np.int64)
    K = 1 << gen
    ngrid = (K + 1) ** d
    grid = ids < ngrid
    r = np.where(grid, ids, ids - ngrid)
    p = np.empty(ids.shape + (d,), np.int64)
    for i in range(d):
        base = np.where(grid, K + 1, K)
        p[..., i] = np.where(grid, 2 * (r % base), 2 * (r % base) + 1)
        r = r // base
    return p


def insertion_key(ids, d, gen):
    """Sort key that lists vertices the way the reference's vertex cache holds them, as far as that
    has a closed form.  The cache is append-only, so a vertex created in an earlier generation
    precedes every vertex of a later one; generation 0 is created as origin, supremum, the other
    corners in binary counting order with axis 0 fastest, centroid last (_complex.py:192-357).
    Within a later generation the reference's order is that of its nested generators and has no
    closed form: ids are used there.  Returns (first generation, secondary key), both int64."""
    p = decode(ids, d, gen)                       # lattice indices at resolution L = 2^(gen+1)
    L = 1 << (gen + 1)
    # trailing zeros of every coordinate (index 0 counts as "as many as needed")
    tz = np.full(p.shape, gen + 2, np.int64)
    nz = p != 0
    tz[nz] = np.log2(p[nz] & -p[nz]).astype(np.int64)
    m = tz.min(axis=-1)
    same = np.all(tz == m[..., None], axis=-1)    # odd multiples of 2^m on every axis: a centroid of generation gen - m
    first = np.where(same, gen - m, gen + 1 - m)
    first = np.clip(first, 0, gen)
    ids = np.asarray(ids, np.int64)
    sec = ids.copy()
    g0 = first == 0
    if g0.any():
        q = p[g0]
        corner = np.all((q == 0) | (q == L), axis=-1)
        bits = (q == L).astype(np.int64)
        count = (bits << np.arange(d, dtype=np.int64)).sum(-1)
        key = count + 1                           # binary counting, axis 0 fastest
        key[np.all(q == 0, axis=-1)] = 0          # origin
        key[np.all(q == L, axis=-1)] = 1          # supremum
        key[~corner] = (1 << d) + 2               # the centroid
        sec[g0] = key
    return first, sec


def neighbours(vid, d, gen):
    """Neighbour ids of one vertex by the parity rule (used only to materialise the few vertex
    objects the driver touches: pool members and their stars)."""
    K = 1 << gen
    ngrid = (K + 1) ** d
    gmul = (K + 1) ** np.arange(d, dtype=np.int64)
    cmul = K ** np.arange(d, dtype=np.int64)
    bits = ((np.arange(1 << d)[:, None] >> np.arange(d)[None, :]) & 1).astype(np.int64)
    if vid >= ngrid:
        r = vid - ngrid
        c = np.array([(r // K ** i) % K for i in range(d)], np.int64)
        return ((c[None, :] + bits) * gmul).sum(1)
    k = np.array([(vid // (K + 1) ** i) % (K + 1) for i in range(d)], np.int64)
    out = []
    cells = k[None, :] - bits
    ok = np.all((cells >= 0) & (cells < K), axis=1)
    out.append(ngrid + (cells[ok] * cmul).sum(1))
    for cls in (1, 0):
        ax = np.flatnonzero((k & 1) == cls)
        if len(ax) == 0:
            continue
        t = np.arange(1, 3 ** len(ax), dtype=np.int64)
        dg = (t[:, None] // 3 ** np.arange(len(ax))[None, :]) % 3
        step = np.where(dg == 1, 1, np.where(dg == 2, -1, 0))
        kk = np.repeat(k[None, :], len(t), 0)
        kk[:, ax] += step
        ok = np.all((kk >= 0) & (kk <= K), axis=1) & ((dg != 0).sum(1) < d)
        out.append((kk[ok] * gmul).sum(1))
    return np.concatenate(out)
