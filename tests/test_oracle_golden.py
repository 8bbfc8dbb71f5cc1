"""Pin the CPU oracle (oracle/) against fixtures recorded from the live reference.

CPU only.  This is what makes the oracle trustworthy as the checker of the CUDA path.
"""
import numpy as np
import pytest

from conftest import (BIT_EXACT_F, LATTICE_CASES, LATTICE_CASES_HIGH_D, SIMPLICIAL_POOL_CASES,
                      SIMPLICIAL_POOL_CASES_HIGH_D, SOBOL_CASES, gset_of,
                      load_golden)
from oracle import oracle as orc


def test_dirnums_match_scipy():
    from scipy.stats import qmc
    for d in (1, 2, 6, 8, 10, 41, 200, 1024):
        sv = orc.sobol_dirnums(d)
        ref = qmc.Sobol(d=d, scramble=False, seed=0)._sv
        assert np.array_equal(sv, np.asarray(ref, np.uint32)), d


def test_sobol_points_match_scipy_with_continuation():
    from scipy.stats import qmc
    for d in (1, 4, 6, 8, 10, 41):
        eng = qmc.Sobol(d=d, scramble=False, seed=0)
        a = eng.random(64)
        b = eng.random(192)          # engine state persists (_shgo.py:1441-1444)
        sv = orc.sobol_dirnums(d)
        lb, ub = np.zeros(d), np.ones(d)
        assert np.array_equal(orc.sobol_points(sv, 0, 64, lb, ub), a)
        assert np.array_equal(orc.sobol_points(sv, 64, 192, lb, ub), b)


@pytest.mark.parametrize("name", SOBOL_CASES)
def test_sobol_case(name):
    g = load_golden(name)
    functor, gset = str(g["functor"]), gset_of(g)
    bounds, n = g["bounds"], int(g["n"])
    d = bounds.shape[0]
    pts = g["points"]
    # a1/a2: points bit-equal to what the reference handed to qhull
    sv = orc.sobol_dirnums(d)
    x = orc.sobol_points(sv, 0, n, bounds[:, 0], bounds[:, 1])
    assert np.array_equal(x, pts)
    # a5: vertex graph from the same simplices
    row_ptr, col = orc.vf_to_vv(g["simplices"].astype(np.int32), n)
    u = np.repeat(np.arange(n), np.diff(row_ptr))
    mine = np.stack([u, col], 1)
    mine = mine[mine[:, 0] < mine[:, 1]]
    assert np.array_equal(mine, g["edges"].astype(np.int64))
    present = (np.diff(row_ptr) > 0).astype(np.uint8)
    assert np.array_equal(present, g["present"])
    # a3/a4: f and feasibility on the vertices that exist
    f, feas = orc.eval_points(functor, gset, x)
    ff, ffe = orc.sobol_eval(sv, functor, gset, 0, n, bounds[:, 0], bounds[:, 1])
    assert np.array_equal(f, ff) and np.array_equal(feas, ffe)
    m = present.astype(bool)
    assert np.array_equal(feas[m], g["feasible"][m])
    fin = m & np.isfinite(g["f"])
    assert np.array_equal(np.isinf(f[m]), np.isinf(g["f"][m]))
    if functor in BIT_EXACT_F:
        assert np.array_equal(f[fin], g["f"][fin])
    else:
        # acceptance criterion: 1e-12 relative, floor = 4 ulp of the largest term (tests/tolerance.py)
        import tolerance
        tolerance.assert_close(f[fin], g["f"][fin], functor, x[fin])
    # a6/a7: minimiser pool and nfev
    pool = orc.compact_pool(orc.star_csr(row_ptr, col, f))
    assert np.array_equal(pool, g["pool"])
    assert orc.nfev(row_ptr, feas) == int(g["nfev"])


@pytest.mark.parametrize("name", LATTICE_CASES + LATTICE_CASES_HIGH_D)
def test_lattice_case(name):
    g = load_golden(name)
    bounds = [tuple(b) for b in g["bounds"]]
    d = len(bounds)
    for gen in range(int(g["gens"]) + 1):
        coords = orc.lattice_coords(bounds, gen)
        ids = {tuple(c): i for i, c in enumerate(coords)}
        assert len(ids) == len(coords)
        ref = g[f"coords_g{gen}"]
        assert len(ref) == len(coords)
        mp = np.array([ids[tuple(c)] for c in ref])        # exact coordinate match
        e = mp[g[f"edges_g{gen}"].astype(np.int64)]
        e = np.sort(e, axis=1)
        e = e[np.lexsort((e[:, 1], e[:, 0]))]
        row_ptr, col = orc.lattice_csr(d, gen)
        u = np.repeat(np.arange(len(coords)), np.diff(row_ptr))
        mine = np.stack([u, col], 1)
        mine = mine[mine[:, 0] < mine[:, 1]]
        assert np.array_equal(mine, e), (name, gen)


def test_lattice_float_artefact_is_detectable():
    """(-3.3, 7.1): the reference itself creates near-duplicate vertices (SURVEY hard part 6)."""
    g = load_golden("lattice_mixed_3d")
    assert len(g["coords_g1"]) == 40 and orc.lattice_nvert(3, 1)[0] == 35


@pytest.mark.parametrize("name", SIMPLICIAL_POOL_CASES + SIMPLICIAL_POOL_CASES_HIGH_D)
def test_simplicial_pools(name):
    """pool_k = strict star minima of generation k minus every earlier local-search start
    vertex (the star() self-loop quirk, _shgo.py:1027-1032 / SURVEY row a6)."""
    g = load_golden(name)
    functor, gset = str(g["functor"]), gset_of(g)
    bounds = [tuple(b) for b in g["bounds"]]
    d = len(bounds)
    started = set()
    for it in range(int(g["iters"])):
        x = orc.lattice_coords(bounds, it)
        assert len(x) == int(g[f"nvert_{it}"])
        f, feas = orc.eval_points(functor, gset, x)
        assert int(feas.sum()) == int(g[f"nfev_{it}"])
        is_min = orc.lattice_star(d, it, f)
        row_ptr, col = orc.lattice_csr(d, it)
        assert np.array_equal(is_min, orc.star_csr(row_ptr, col, f))
        pool = {tuple(x[i]) for i in np.flatnonzero(is_min)} - started
        ref = {tuple(p) for p in g[f"pool_x_{it}"]}
        assert pool == ref, (name, it)
        started |= ref


def local_bounds_names():
    return [str(s) for s in load_golden("local_bounds")["names"]]


@pytest.mark.parametrize("name", local_bounds_names())
def test_local_bounds(name):
    """f2: construct_lcb_simplicial of every vertex (recorded through the reference's method)."""
    g = load_golden("local_bounds")
    x, edges, bounds = g[name + "_x"], g[name + "_edges"], g[name + "_bounds"]
    row_ptr, col = orc.csr_from_edges(edges, len(x))
    lo, hi = orc.local_bounds_csr(row_ptr, col, x, np.arange(len(x)), bounds)
    assert np.array_equal(lo, g[name + "_lo"])
    assert np.array_equal(hi, g[name + "_hi"])
