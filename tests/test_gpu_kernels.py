"""Parity of the CUDA hot path against the CPU oracle and the golden fixtures (B200 only).

Every call goes through the C ABI (shgo_b200.hotpath is a ctypes shim over libshgo_b200.so).
Bars: Sobol points, adjacency, feasibility and pool index sets bit-exact; f bit-exact for the
polynomial / FMA-chain functors and within 1e-12 RELATIVE for the ones that go through cos/sin/exp,
with the documented absolute floor of tests/tolerance.py (4 ulp of the largest term of the sum).
"""
import numpy as np
import pytest

from conftest import (BIT_EXACT_F, LATTICE_CASES, SIMPLICIAL_POOL_CASES, SOBOL_CASES, gset_of,
                      load_golden)

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import oracle as orc  # noqa: E402

import tolerance  # noqa: E402

DEVICE_BIT_EXACT = BIT_EXACT_F | {"styblinski_tang"}   # device == C oracle bit for bit


@pytest.fixture(scope="module")
def hp():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from shgo_b200 import hotpath
    return hotpath


def _ids(functor, gset):
    return orc.FUNCTOR_IDS[functor], orc.GSET_IDS[gset]


def _close(a, b, functor, x):
    """north_star's bar: 1e-12 relative (+ the explicit floor, tests/tolerance.py); x = (n, d) points"""
    tolerance.assert_close(a, b, functor, x)


# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d,k0,n", [(1, 0, 1), (2, 0, 1000), (6, 0, 4096), (8, 12345, 70001),
                                    (10, 1 << 20, 5000), (32, 7, 3000), (4, (1 << 30) - 4096, 4096)])
def test_sobol_points_bit_exact(hp, d, k0, n):
    rng = np.random.default_rng(d)
    lb = rng.uniform(-50, 0, d)
    ub = lb + rng.uniform(0.5, 80, d)
    if d == 6:
        lb[:], ub[:] = -5.12, 5.12
    tab = hp.SobolTables(np.stack([lb, ub], 1))
    ref = orc.sobol_points(orc.sobol_dirnums(d), k0, n, lb, ub)
    soa = hp.sobol_points(tab, k0, n, "soa").cpu().numpy()
    aos = hp.sobol_points(tab, k0, n, "aos").cpu().numpy()
    assert np.array_equal(soa.T, ref)
    assert np.array_equal(aos, ref)


def test_sobol_points_many_axes(hp):
    """41 axes (reference tests/test__shgo.py:719-728 `test_13_high_sobol`) and 100 axes"""
    from scipy.stats import qmc
    for d in (41, 100):
        tab = hp.SobolTables([(0.0, 1.0)] * d)
        ref = qmc.Sobol(d=d, scramble=False, seed=0).random(256)
        assert np.array_equal(hp.sobol_points(tab, 0, 256, "aos").cpu().numpy(), ref)
        assert np.array_equal(hp.sobol_points(tab, 0, 256, "soa").cpu().numpy().T, ref)


def test_csr_rejects_out_of_range_simplices(hp):
    simp = torch.tensor([[0, 1, 2], [1, 2, 7]], dtype=torch.int32, device="cuda")
    with pytest.raises(ValueError):
        hp.csr_from_simplices(simp, 5)


def test_sobol_points_match_scipy(hp):
    from scipy.stats import qmc
    d, n = 6, 2048
    tab = hp.SobolTables([(-5.12, 5.12)] * d)
    C = qmc.Sobol(d=d, scramble=False, seed=0).random(n)
    for i in range(d):                      # the reference's scaling loop, _shgo.py:1446-1449
        C[:, i] = C[:, i] * (5.12 - -5.12) + -5.12
    assert np.array_equal(hp.sobol_points(tab, 0, n, "aos").cpu().numpy(), C)


def test_empty_and_limits(hp):
    tab = hp.SobolTables([(0, 1)] * 3)
    assert hp.sobol_points(tab, 5, 0).shape == (3, 0)
    f, feas = hp.eval_sobol(3, 0, tab, 0, 0)
    assert f.numel() == 0
    from shgo_b200 import _lib
    with pytest.raises(_lib.Shgob200Error):
        hp.sobol_points(tab, (1 << 30) - 10, 11)       # past the 30-bit sequence
    with pytest.raises(_lib.Shgob200Error):
        hp.eval_sobol(4, 0, tab, 0, 16)                # eggholder is 2-D only
    with pytest.raises(_lib.Shgob200Error):
        hp.eval_sobol(1, 1, tab, 0, 16)                # unregistered functor/constraint pair


# ------------------------------------------------------------------------------------------
EVAL_CASES = [
    ("rosenbrock", None, [(0, 2)] * 5), ("rosenbrock", None, [(-2, 2)] * 12),
    ("rastrigin", None, [(-5.12, 5.12)] * 6), ("rastrigin", None, [(-5.12, 5.12)] * 11),
    ("ackley", None, [(-32.768, 32.768)] * 10), ("ackley", None, [(-32.768, 32.768)] * 3),
    ("styblinski_tang", None, [(-5, 5)] * 8), ("styblinski_tang", None, [(-5, 5)] * 3),
    ("eggholder", None, [(-512, 512)] * 2),
    ("cattle_feed", None, [(0, 1.0)] * 4), ("cattle_feed", "cattle_feed", [(0, 1.0)] * 4),
    ("sphere", None, [(-1, 6)] * 4), ("sphere", "sum_le_6", [(-1, 6)] * 2),
    # one axis: torch gives a (1, n) tensor an arbitrary stride(0) (round-1 failure `ld < n`)
    ("sphere", None, [(-1, 6)]), ("rastrigin", None, [(-2.1, 4.4)]), ("styblinski_tang", None, [(-2.0, 6.2)]),
    ("ackley", None, [(-3.0, 5.0)]),
]


@pytest.mark.parametrize("functor,gset,bounds", EVAL_CASES)
def test_eval_sobol_vs_oracle(hp, functor, gset, bounds):
    fid, gid = _ids(functor, gset)
    b = np.array(bounds, float)
    d = len(bounds)
    tab = hp.SobolTables(bounds)
    sv = orc.sobol_dirnums(d)
    for k0, n in [(0, 20000), (777, 4097)]:
        f, feas = hp.eval_sobol(fid, gid, tab, k0, n)
        fo, feo = orc.sobol_eval(sv, functor, gset, k0, n, b[:, 0], b[:, 1])
        f, feas = f.cpu().numpy(), feas.cpu().numpy()
        assert np.array_equal(feas, feo)
        if functor in DEVICE_BIT_EXACT:
            assert np.array_equal(f, fo)
        else:
            _close(f, fo, functor, orc.sobol_points(sv, k0, n, b[:, 0], b[:, 1]))
        # same evaluation on materialised points
        x = hp.sobol_points(tab, k0, n, "soa")
        f2, feas2 = hp.eval_points(fid, gid, x)
        assert torch.equal(f2.cpu(), torch.from_numpy(f)) and np.array_equal(feas2.cpu().numpy(), feas)


def test_eval_matches_numpy_callables(hp):
    """device f vs the reference-side numpy callables (what compute_sfield would call)."""
    from shgo_b200 import functors as F
    for fn, bounds in [(F.rosenbrock, [(0, 2)] * 5), (F.rastrigin, [(-5.12, 5.12)] * 6),
                       (F.ackley, [(-32.768, 32.768)] * 10), (F.styblinski_tang, [(-5, 5)] * 8),
                       (F.eggholder, [(-512, 512)] * 2), (F.cattle_feed, [(0, 1.0)] * 4)]:
        tab = hp.SobolTables(bounds)
        n = 2000
        x = hp.sobol_points(tab, 0, n, "aos").cpu().numpy()
        f = hp.eval_sobol(fn.shgo_b200_functor, 0, tab, 0, n)[0].cpu().numpy()
        ref = np.array([fn(xi) for xi in x])
        if fn in (F.rosenbrock, F.cattle_feed):
            assert np.array_equal(f, ref)
        else:
            _close(f, ref, fn.__name__, x)


@pytest.mark.parametrize("functor,d,lo,hi", [
    ("rastrigin", 1, -2.1, 4.4), ("styblinski_tang", 1, -2.0, 6.2), ("sphere", 1, -1.0, 6.0),
    ("ackley", 1, -3.0, 5.0), ("rastrigin", 6, -5.12, 5.12), ("ackley", 10, -32.768, 32.768),
    ("eggholder", 2, -512.0, 512.0), ("rosenbrock", 5, 0.0, 2.0),
])
def test_eval_points_on_non_sobol_points(hp, functor, d, lo, hi):
    """Halton / custom-sampler / second-iteration route: the host hands over an (n, d) array, the
    complex uploads it as `torch.from_numpy(pts).cuda().t().contiguous()` (complex.py) -- for d = 1
    that tensor has stride(0) == 1.  Every registered functor, odd n, also through a strided view."""
    rng = np.random.default_rng(d * 7 + len(functor))
    for n in (1, 94, 4097):
        pts = rng.uniform(lo, hi, size=(n, d))
        x = torch.from_numpy(pts).cuda().t().contiguous()
        f, feas = hp.eval_points(orc.FUNCTOR_IDS[functor], 0, x)
        fo, feo = orc.eval_points(functor, None, pts)
        assert np.array_equal(feas.cpu().numpy(), feo)
        if functor in DEVICE_BIT_EXACT:
            assert np.array_equal(f.cpu().numpy(), fo)
        else:
            _close(f.cpu().numpy(), fo, functor, pts)
        wide = torch.zeros((d, n + 13), dtype=torch.float64, device="cuda")     # ld > n
        wide[:, :n] = x
        f2, _ = hp.eval_points(orc.FUNCTOR_IDS[functor], 0, wide[:, :n])
        assert torch.equal(f2, f)


@pytest.mark.parametrize("functor,d,width", [("ackley", 10, 32.768), ("ackley", 3, 32.768), ("rastrigin", 6, 5.12),
                                             ("rastrigin", 2, 5.12)])
def test_transcendental_functors_near_their_minimum(hp, functor, d, width):
    """Where an absolute tolerance would hide everything: f -> 0 at the origin (Ackley: the sum passes
    through 20 + e; Rastrigin: through 10 d).  Points at distances 1e-1 ... 1e-9 from the minimiser,
    the minimiser itself, and the neighbouring integer-lattice minima of Rastrigin."""
    rng = np.random.default_rng(5)
    pts = [np.zeros((1, d))]
    for r in 10.0 ** -np.arange(1, 10):
        pts.append(rng.standard_normal((400, d)) * r)
    if functor == "rastrigin":
        pts.append(np.round(rng.uniform(-4, 4, (400, d))) + rng.standard_normal((400, d)) * 1e-4)
    pts = np.clip(np.concatenate(pts), -width, width)
    x = torch.from_numpy(pts).cuda().t().contiguous()
    f, _ = hp.eval_points(orc.FUNCTOR_IDS[functor], 0, x)
    from shgo_b200 import functors as F
    ref = np.array([getattr(F, functor)(p) for p in pts])         # the reference-side numpy callable
    _close(f.cpu().numpy(), ref, functor, pts)
    _close(f.cpu().numpy(), orc.eval_points(functor, None, pts)[0], functor, pts)


def test_eval_sobol_block_minima(hp):
    """group minima of the outputs (+inf for infeasible / NaN) in the layout of include/shgo_b200.h"""
    for fname, gname, bounds in (("cattle_feed", "cattle_feed", [(0.0, 1.0)] * 4), ("styblinski_tang", None, [(-5.0, 5.0)] * 8),
                                 ("ackley", None, [(-32.768, 32.768)] * 3)):
        fid, gid = orc.FUNCTOR_IDS[fname], orc.GSET_IDS[gname]
        plog = hp.blockmin_layout(fid, gid)
        tile = 256 << plog
        tab = hp.SobolTables(bounds)
        ntot = 8 * tile
        bm = torch.full((ntot // 32,), -7.0, dtype=torch.float64, device="cuda")
        for k0, n in ((0, 3 * tile), (3 * tile, 5 * tile)):
            f = torch.empty(n, dtype=torch.float64, device="cuda")
            feas = torch.empty(n, dtype=torch.uint8, device="cuda")
            hp.eval_sobol_blockmin(fid, gid, tab, k0, n, f, feas, bm)
            f2, feas2 = hp.eval_sobol(fid, gid, tab, k0, n)
            assert torch.equal(f, f2) and torch.equal(feas, feas2)
        fall, _ = hp.eval_sobol(fid, gid, tab, 0, ntot)
        ref = np.full(ntot // 32, np.inf)
        np.minimum.at(ref, hp.blockmin_group(np.arange(ntot), plog), fall.cpu().numpy())
        assert np.array_equal(bm.cpu().numpy(), ref), fname
        from shgo_b200 import _lib
        with pytest.raises(_lib.Shgob200Error):
            hp.eval_sobol_blockmin(fid, gid, tab, 32, tile, f, feas, bm)
    assert hp.blockmin_layout(1, 1) == -1


def test_mask_infeasible(hp):
    n = 10000
    rng = np.random.default_rng(3)
    f = rng.standard_normal(n)
    f[::17] = np.nan
    g = rng.standard_normal((2, n))
    g[0, ::13] = np.nan                      # NaN constraint value counts as feasible
    ft, feas = hp.mask_infeasible(torch.from_numpy(f).cuda(), torch.from_numpy(g).cuda())
    ok = ~((g[0] < 0) | (g[1] < 0))
    ref = np.where(ok & ~np.isnan(f), f, np.inf)
    assert np.array_equal(ft.cpu().numpy(), ref)
    assert np.array_equal(feas.cpu().numpy().astype(bool), ok)


# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", SOBOL_CASES)
def test_golden_sobol_case(hp, name):
    g = load_golden(name)
    functor, gset = str(g["functor"]), gset_of(g)
    fid, gid = _ids(functor, gset)
    bounds, n = g["bounds"], int(g["n"])
    tab = hp.SobolTables(bounds)
    # points handed to qhull
    assert np.array_equal(hp.sobol_points(tab, 0, n, "aos").cpu().numpy(), g["points"])
    # vf_to_vv on the very simplices the reference consumed
    simp = torch.from_numpy(g["simplices"].astype(np.int32)).cuda()
    row_ptr, col = hp.csr_from_simplices(simp, n)
    rp_o, col_o = orc.vf_to_vv(g["simplices"].astype(np.int32), n)
    assert np.array_equal(row_ptr.cpu().numpy(), rp_o)
    assert np.array_equal(col.cpu().numpy(), col_o)
    rp = row_ptr.cpu().numpy()
    u = np.repeat(np.arange(n), np.diff(rp))
    e = np.stack([u, col.cpu().numpy()], 1)
    assert np.array_equal(e[e[:, 0] < e[:, 1]], g["edges"].astype(np.int64))
    # evaluate + mask
    f, feas = hp.eval_sobol(fid, gid, tab, 0, n)
    present = np.diff(rp) > 0
    fh = f.cpu().numpy()
    assert np.array_equal(feas.cpu().numpy()[present], g["feasible"][present])
    if functor in BIT_EXACT_F:
        assert np.array_equal(fh[present], g["f"][present])
    else:
        _close(fh[present], g["f"][present], functor, g["points"][present])
    # classify + pool + bookkeeping
    is_min = hp.star_csr(row_ptr, col, f)
    idx, cnt = hp.compact_pool(is_min)
    pool = idx[: int(cnt.item())].cpu().numpy()
    assert np.array_equal(pool, g["pool"])
    assert int(hp.count_nfev(row_ptr, feas).item()) == int(g["nfev"])
    nf = torch.zeros(1, dtype=torch.int64, device="cuda")       # fused star + nfev pass
    is_min2 = hp.star_csr(row_ptr, col, f, feasible=feas, nfev=nf)
    assert torch.equal(is_min2, is_min) and int(nf.item()) == int(g["nfev"])
    fo, _ = orc.eval_points(functor, gset, g["points"])
    pres_t = torch.from_numpy(present.astype(np.uint8)).cuda()
    i_dev, v_dev = hp.argmin(f, pres_t)
    i_or = orc.argmin(fh, present.astype(np.uint8))
    assert i_dev == i_or and v_dev == fh[i_or]


def test_star_sharded_rows_and_random_graphs(hp):
    rng = np.random.default_rng(11)
    for n, S, dp1 in [(1000, 5000, 4), (50, 7, 3), (4096, 60000, 7), (300, 299, 2), (10, 0, 3)]:
        simp = rng.integers(0, n, size=(S, dp1)).astype(np.int32)
        rp_o, col_o = orc.vf_to_vv(simp, n)
        row_ptr, col = hp.csr_from_simplices(torch.from_numpy(simp).cuda(), n)
        assert np.array_equal(row_ptr.cpu().numpy(), rp_o)
        assert np.array_equal(col.cpu().numpy(), col_o)
        f = rng.standard_normal(n)
        f[rng.integers(0, n, n // 10)] = np.inf
        f[rng.integers(0, n, n // 10)] = f[rng.integers(0, n, n // 10)]     # exact ties
        ft = torch.from_numpy(f).cuda()
        ref = orc.star_csr(rp_o, col_o, f)
        assert np.array_equal(hp.star_csr(row_ptr, col, ft).cpu().numpy(), ref)
        for row0, rows in [(0, n // 3), (n // 3, n - n // 3), (n - 1, 1), (5, 0)]:
            got = hp.star_csr(row_ptr, col, ft, row0, rows).cpu().numpy()
            assert np.array_equal(got, ref[row0:row0 + rows])
        idx, cnt = hp.compact_pool(torch.from_numpy(ref).cuda(), base=100)
        assert np.array_equal(idx[: int(cnt.item())].cpu().numpy(), orc.compact_pool(ref) + 100)


def test_star_skewed_degrees_fall_back_to_global_col(hp):
    """one hub with ~20000 neighbours: the tile's slice exceeds the staging buffer"""
    n = 30000
    hub = np.stack([np.zeros(n - 1, np.int32), np.arange(1, n, dtype=np.int32)], 1)
    chain = np.stack([np.arange(1, n - 1, dtype=np.int32), np.arange(2, n, dtype=np.int32)], 1)
    simp = np.concatenate([hub, chain])
    rng = np.random.default_rng(5)
    f = rng.standard_normal(n)
    f[0] = -100.0
    rp_o, col_o = orc.vf_to_vv(simp, n)
    row_ptr, col = hp.csr_from_simplices(torch.from_numpy(simp).cuda(), n)
    assert np.array_equal(col.cpu().numpy(), col_o)
    got = hp.star_csr(row_ptr, col, torch.from_numpy(f).cuda()).cpu().numpy()
    assert np.array_equal(got, orc.star_csr(rp_o, col_o, f)) and got[0] == 1


def test_compact_and_argmin_edge_cases(hp):
    z = torch.zeros(5000, dtype=torch.uint8, device="cuda")
    idx, cnt = hp.compact_pool(z)
    assert int(cnt.item()) == 0
    o = torch.ones(4097, dtype=torch.uint8, device="cuda")
    idx, cnt = hp.compact_pool(o, cap=10)
    assert int(cnt.item()) == 4097 and idx[:10].tolist() == list(range(10))
    f = torch.full((1000,), float("inf"), dtype=torch.float64, device="cuda")
    assert hp.argmin(f) == (-1, float("inf"))
    f[700] = 3.0
    f[300] = 3.0
    assert hp.argmin(f) == (300, 3.0)


# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", LATTICE_CASES)
def test_lattice_coords_and_star_vs_oracle(hp, name):
    g = load_golden(name)
    bounds = [tuple(b) for b in g["bounds"]]
    d = len(bounds)
    from shgo_b200 import lattice
    rng = np.random.default_rng(1)
    for gen in range(int(g["gens"]) + 1):
        tab_h = lattice.coordinate_tables(bounds, gen)
        assert np.array_equal(tab_h, orc.lattice_tables(bounds, gen))
        tab = torch.from_numpy(tab_h).cuda()
        nv, _ = orc.lattice_nvert(d, gen)
        x = hp.lattice_coords(d, gen, tab, 0, nv).cpu().numpy().T
        assert np.array_equal(x, orc.lattice_coords(bounds, gen))
        assert {tuple(c) for c in x} == {tuple(c) for c in g[f"coords_g{gen}"]}
        for trial in range(3):
            f = rng.standard_normal(nv) if trial else np.sum(x ** 2, axis=1)
            if trial == 2:
                f[rng.integers(0, nv, nv // 5 + 1)] = np.inf
            got = hp.lattice_star(d, gen, torch.from_numpy(f).cuda()).cpu().numpy()
            assert np.array_equal(got, orc.lattice_star(d, gen, f)), (name, gen, trial)


@pytest.mark.parametrize("name", SIMPLICIAL_POOL_CASES)
def test_golden_simplicial_pools(hp, name):
    g = load_golden(name)
    functor, gset = str(g["functor"]), gset_of(g)
    fid, gid = _ids(functor, gset)
    bounds = [tuple(b) for b in g["bounds"]]
    d = len(bounds)
    from shgo_b200 import lattice
    started = set()
    for it in range(int(g["iters"])):
        tab = torch.from_numpy(lattice.coordinate_tables(bounds, it)).cuda()
        nv, _ = orc.lattice_nvert(d, it)
        assert nv == int(g[f"nvert_{it}"])
        f, feas = hp.lattice_eval(fid, gid, d, it, tab, 0, nv)
        assert int(feas.sum().item()) == int(g[f"nfev_{it}"])
        x = hp.lattice_coords(d, it, tab, 0, nv).cpu().numpy().T
        fo, feo = orc.eval_points(functor, gset, x)
        if functor in DEVICE_BIT_EXACT:
            assert np.array_equal(f.cpu().numpy(), fo)
        else:
            _close(f.cpu().numpy(), fo, functor, x)
        is_min = hp.lattice_star(d, it, f).cpu().numpy()
        pool = {tuple(x[i]) for i in np.flatnonzero(is_min)} - started
        ref = {tuple(p) for p in g[f"pool_x_{it}"]}
        assert pool == ref, (name, it)
        started |= ref


def test_lattice_10d_generation1_vs_oracle(hp):
    """config 3 scale-down: 10-D Ackley lattice, generation 1 (60 073 vertices, degree <= 59 048)"""
    d, gen = 10, 1
    bounds = [(-32.768, 32.768)] * d
    from shgo_b200 import lattice
    tab = torch.from_numpy(lattice.coordinate_tables(bounds, gen)).cuda()
    nv, _ = orc.lattice_nvert(d, gen)
    f, _ = hp.lattice_eval(orc.FUNCTOR_IDS["ackley"], 0, d, gen, tab, 0, nv)
    fh = f.cpu().numpy()
    x = orc.lattice_coords(bounds, gen)
    _close(fh, orc.eval_points("ackley", None, x)[0], "ackley", x)
    got = hp.lattice_star(d, gen, f).cpu().numpy()
    assert np.array_equal(got, orc.lattice_star(d, gen, fh))
    # sharded vertex ranges give the same flags
    a = hp.lattice_star(d, gen, f, 0, 30000).cpu().numpy()
    b = hp.lattice_star(d, gen, f, 30000, nv - 30000).cpu().numpy()
    assert np.array_equal(np.concatenate([a, b]), got)


# ------------------------------------------------------------------------------------------
def test_full_size_properties(hp):
    """BASELINE sizes (config 5 at 2^24): properties that do not need an O(n) CPU pass in Python,
    plus oracle spot checks on shards."""
    n, d = 1 << 24, 8
    bounds = [(-5.0, 5.0)] * d
    tab = hp.SobolTables(bounds)
    f, feas = hp.eval_sobol(orc.FUNCTOR_IDS["styblinski_tang"], 0, tab, 0, n)
    sv = orc.sobol_dirnums(d)
    b = np.array(bounds)
    for k0 in (0, 5_000_001, n - 65536):
        fo, _ = orc.sobol_eval(sv, "styblinski_tang", None, k0, 65536, b[:, 0], b[:, 1])
        assert np.array_equal(f[k0:k0 + 65536].cpu().numpy(), fo)
    # shard-additivity: evaluating [k0, k0+m) alone equals the slice of the whole
    f2, _ = hp.eval_sobol(orc.FUNCTOR_IDS["styblinski_tang"], 0, tab, 3_333_333, 1_000_000)
    assert torch.equal(f2, f[3_333_333:4_333_333])
    row_ptr, col = hp.hypercube_csr(n)
    is_min = hp.star_csr(row_ptr, col, f)
    # definition check on the device: min over neighbours, strict
    m = 24
    nbmin = torch.full((n,), float("inf"), dtype=torch.float64, device="cuda")
    i = torch.arange(n, device="cuda")
    for j in range(m):
        nbmin = torch.minimum(nbmin, f[i ^ (1 << j)])
    assert torch.equal(is_min.bool(), f < nbmin)
    idx, cnt = hp.compact_pool(is_min)
    c = int(cnt.item())
    assert c == int(is_min.sum().item())
    pool = idx[:c]
    assert bool((pool[1:] > pool[:-1]).all()) and bool(is_min[pool].all())
    # oracle on a row range of the same graph
    rp_h, col_h = orc.hypercube_csr(1 << 16)
    fh = f[: 1 << 16].cpu().numpy()
    sub = hp.star_csr(torch.from_numpy(rp_h).cuda(), torch.from_numpy(col_h).cuda(),
                      f[: 1 << 16].contiguous()).cpu().numpy()
    assert np.array_equal(sub, orc.star_csr(rp_h, col_h, fh))


def test_host_buffer_iteration_matches_device_path(hp):
    import ctypes
    from shgo_b200 import _lib
    n, d = 1 << 16, 6
    bounds = np.array([(-5.12, 5.12)] * d)
    sv = orc.sobol_dirnums(d)
    rp, col = orc.hypercube_csr(n)
    fo, feo = orc.sobol_eval(sv, "rastrigin", None, 0, n, bounds[:, 0], bounds[:, 1])
    ctx = ctypes.c_void_p()
    _lib.call("shgo_b200_ctx_create", 0, ctypes.byref(ctx))
    try:
        pool = np.empty(n, np.int64)
        cnt, nfev = ctypes.c_uint64(), ctypes.c_uint64()
        fout = np.empty(n)
        lb, ub = bounds[:, 0].copy(), bounds[:, 1].copy()
        for _ in range(2):
            _lib.call("shgo_b200_iterate_sobol_host", ctx, 1, 0, d, 0, n, sv.ctypes.data, lb.ctypes.data,
                      ub.ctypes.data, rp.ctypes.data, col.ctypes.data, pool.ctypes.data, n,
                      ctypes.byref(cnt), ctypes.byref(nfev), fout.ctypes.data)
        _close(fout, fo, "rastrigin", orc.sobol_points(sv, 0, n, lb, ub))
        ref_pool = orc.compact_pool(orc.star_csr(rp, col, fout))
        assert np.array_equal(pool[: cnt.value], ref_pool)
        assert nfev.value == n
    finally:
        _lib.lib().shgo_b200_ctx_destroy(ctx)


@pytest.mark.parametrize("nshards", [2, 8])
def test_two_phase_sharded_star_emulated_on_one_gpu(hp, nshards):
    """The multi-GPU star test (local phase + remote phase over peer pointers) with all `shards`
    living on this one GPU: every shard is classified in turn and the union must equal the
    single-GPU result.  (The real multi-process path is tools/check_sharded.py under torchrun.)"""
    rng = np.random.default_rng(7)
    for graph in ("hypercube", "random"):
        n = 1 << 14
        if graph == "hypercube":
            rp_h, col_h = orc.hypercube_csr(n)
        else:
            simp = rng.integers(0, n, size=(6 * n, 4)).astype(np.int32)
            rp_h, col_h = orc.vf_to_vv(simp, n)
        f_h = rng.standard_normal(n)
        f_h[rng.integers(0, n, n // 20)] = np.inf
        ref = orc.star_csr(rp_h, col_h, f_h)
        rows = n // nshards
        f_sh = [torch.from_numpy(f_h[s * rows:(s + 1) * rows].copy()).cuda() for s in range(nshards)]
        table = torch.tensor([t.data_ptr() for t in f_sh], dtype=torch.int64, device="cuda")
        got = np.empty(n, np.uint8)
        ncand = 0
        for s in range(nshards):
            row0 = s * rows
            rp_l = torch.from_numpy(rp_h[row0:row0 + rows + 1].copy()).cuda()
            c = col_h[rp_h[row0]:rp_h[row0 + rows]]
            col_l = torch.zeros(max(len(c), 4), dtype=torch.int32, device="cuda")
            col_l[:len(c)] = torch.from_numpy(c.copy()).cuda()
            col_l = col_l[:len(c)]
            flags = torch.empty(rows, dtype=torch.uint8, device="cuda")
            nfev = torch.zeros(1, dtype=torch.int64, device="cuda")
            hp.star_csr_local(rp_l, col_l, row0, f_sh[s], None, flags, nfev)
            if nshards == 2:       # chunked classification of the same shard gives the same flags
                fl2 = torch.empty_like(flags)
                nf2 = torch.zeros(2, dtype=torch.int64, device="cuda")
                h = rows // 2
                hp.star_csr_local(rp_l, col_l, row0, f_sh[s], None, fl2, nf2[0:1], None, row0, h)
                hp.star_csr_local(rp_l, col_l, row0, f_sh[s], None, fl2, nf2[1:2], None, row0 + h, rows - h)
                assert torch.equal(fl2, flags) and int(nf2.sum().item()) == int(nfev.item())
            fl = flags.cpu().numpy()
            ncand += int((fl == 2).sum())
            # phase-1 flags are already final wherever they are 0, and 1 means "no remote neighbour"
            assert not np.any((fl == 0) & (ref[row0:row0 + rows] == 1))
            ws = __import__("shgo_b200")._lib.lib().shgo_b200_star_remote_workspace(rows, rows)
            tmp = torch.empty(ws, dtype=torch.uint8, device="cuda")
            fl_slow = flags.clone()
            hp.star_csr_remote(rp_l, col_l, row0, table, f_sh[s].data_ptr(), nshards, rows, s, fl_slow, tmp)
            # both test graphs keep remote neighbours at the row ends (top bits last / ascending rows)
            hp.star_csr_remote(rp_l, col_l, row0, table, f_sh[s].data_ptr(), nshards, rows, s, flags, tmp,
                               remote_at_ends=True)
            assert torch.equal(flags, fl_slow)
            # phase 1 listing its candidates itself, phase 2 consuming that list
            from shgo_b200 import _lib as L
            cws = torch.empty(L.lib().shgo_b200_star_local_workspace(rows), dtype=torch.uint8, device="cuda")
            fl3 = torch.empty_like(flags)
            hp.star_csr_local(rp_l, col_l, row0, f_sh[s], None, fl3, nfev, cand_ws=cws)
            hp.star_csr_remote(rp_l, col_l, row0, table, f_sh[s].data_ptr(), nshards, rows, s, fl3, None,
                               remote_at_ends=True, cand_ws=cws)
            assert torch.equal(fl3, flags)
            # phase 1 trimming every row to its local middle part (the same promise)
            fl4 = torch.empty_like(flags)
            nf4 = torch.zeros(1, dtype=torch.int64, device="cuda")
            hp.star_csr_local(rp_l, col_l, row0, f_sh[s], None, fl4, nf4, cand_ws=cws, remote_at_ends=True)
            f4 = fl4.cpu().numpy()
            # identical, except that a row WITHOUT local neighbours is always left to phase 2 (the
            # checked walk already rejects it when its own f is +inf)
            assert np.all((f4 == fl) | ((f4 == 2) & (fl == 0))) and int(nf4.item()) == int(nfev.item())
            hp.star_csr_remote(rp_l, col_l, row0, table, f_sh[s].data_ptr(), nshards, rows, s, fl4, None,
                               remote_at_ends=True, cand_ws=cws)
            assert torch.equal(fl4, flags)
            # phase 2 with the conservative block-minimum filter: same flags, fewer remote reads
            plog = 2
            grp = hp.blockmin_group(np.arange(n), plog)
            bm_h = np.full(n // 32, np.inf)
            np.minimum.at(bm_h, grp, np.where(np.isnan(f_h), np.inf, f_h))
            bmin = torch.from_numpy(bm_h).cuda()
            fl6 = torch.empty_like(flags)
            hp.star_csr_local(rp_l, col_l, row0, f_sh[s], None, fl6, nfev, cand_ws=cws, remote_at_ends=True)
            hp.star_csr_remote(rp_l, col_l, row0, table, f_sh[s].data_ptr(), nshards, rows, s, fl6, None,
                               remote_at_ends=True, cand_ws=cws, block_min=bmin, layout_plog=plog)
            assert torch.equal(fl6, flags)
            fl7 = torch.empty_like(flags)
            hp.star_csr_local(rp_l, col_l, row0, f_sh[s], None, fl7, nfev)
            hp.star_csr_remote(rp_l, col_l, row0, table, f_sh[s].data_ptr(), nshards, rows, s, fl7, tmp,
                               block_min=bmin, layout_plog=plog)
            assert torch.equal(fl7, flags)
            # phases 1 + 2 fused in one kernel: peers all ready / none ready / some ready
            for ready_mask in (None, 0, 0b10101010):
                epoch = 7
                rdy = torch.full((nshards,), epoch, dtype=torch.int32, device="cuda")
                if ready_mask is not None:
                    for q in range(nshards):
                        if not (ready_mask >> q) & 1:
                            rdy[q] = epoch - 1
                fl5 = torch.full_like(flags, 9)
                nf5 = torch.zeros(1, dtype=torch.int64, device="cuda")
                hp.star_csr_fused(rp_l, col_l, row0, table, f_sh[s], None, nshards, rows, s, rdy, epoch, fl5, cws, nf5)
                f5 = fl5.cpu().numpy()
                fin = flags.cpu().numpy()
                assert np.all((f5 == fin) | (f5 == 2)), (graph, nshards, ready_mask)
                if ready_mask is None and graph == "hypercube":
                    # <= 3 remote neighbours per row: everything inline but tiles with > 8 candidates
                    assert np.count_nonzero(f5 == 2) <= np.count_nonzero(f4 == 2) // 50
                if ready_mask == 0:
                    assert np.array_equal(f5 == 2, f4 == 2)
                assert int(nf5.item()) == int(nfev.item())
                hp.star_csr_remote(rp_l, col_l, row0, table, f_sh[s].data_ptr(), nshards, rows, s, fl5, None,
                                   remote_at_ends=True, cand_ws=cws)
                assert torch.equal(fl5, flags), (graph, nshards, ready_mask)
            got[row0:row0 + rows] = flags.cpu().numpy()
            assert int(nfev.item()) == int(np.count_nonzero(np.diff(rp_h[row0:row0 + rows + 1]) > 0))
        assert np.array_equal(got, ref), (graph, nshards)
        assert ncand > 0


@pytest.mark.parametrize("functor,gset,bounds,log2n", [
    ("rastrigin", None, [(-5.12, 5.12)] * 6, 20),               # BASELINE config 2
    ("cattle_feed", "cattle_feed", [(0, 1.0)] * 4, 22),         # BASELINE config 4
])
def test_baseline_configs_at_full_size(hp, functor, gset, bounds, log2n):
    """configs 2 and 4 at their named sizes on the synthetic index-hypercube CSR: oracle spot checks
    of f / feasibility on shards, the star test against its definition, pool and nfev consistency."""
    fid, gid = _ids(functor, gset)
    n, d = 1 << log2n, len(bounds)
    b = np.array(bounds, float)
    tab = hp.SobolTables(bounds)
    f, feas = hp.eval_sobol(fid, gid, tab, 0, n)
    sv = orc.sobol_dirnums(d)
    for k0 in (0, n // 3, n - 8192):
        fo, feo = orc.sobol_eval(sv, functor, gset, k0, 8192, b[:, 0], b[:, 1])
        assert np.array_equal(feas[k0:k0 + 8192].cpu().numpy(), feo)
        if functor in DEVICE_BIT_EXACT:
            assert np.array_equal(f[k0:k0 + 8192].cpu().numpy(), fo)
        else:
            _close(f[k0:k0 + 8192].cpu().numpy(), fo, functor, orc.sobol_points(sv, k0, 8192, b[:, 0], b[:, 1]))
    assert bool(torch.isinf(f[feas == 0]).all()) and bool(torch.isfinite(f[feas == 1]).all())
    row_ptr, col = hp.hypercube_csr(n)
    nf = torch.zeros(1, dtype=torch.int64, device="cuda")
    is_min = hp.star_csr(row_ptr, col, f, feasible=feas, nfev=nf)
    assert int(nf.item()) == int(feas.sum().item())
    nbmin = torch.full((n,), float("inf"), dtype=torch.float64, device="cuda")
    i = torch.arange(n, device="cuda")
    for j in range(log2n):
        nbmin = torch.minimum(nbmin, f[i ^ (1 << j)])
    assert torch.equal(is_min.bool(), f < nbmin)
    assert not bool(is_min[feas == 0].any())                     # an infeasible vertex is never a minimiser
    idx, cnt = hp.compact_pool(is_min)
    pool = idx[: int(cnt.item())]
    assert int(cnt.item()) == int(is_min.sum().item()) and bool((pool[1:] > pool[:-1]).all())
    # the oracle on the first 2^15 rows of the same graph restricted to them is not meaningful (the
    # neighbours leave the range), so compare on a self-contained small instance instead
    m = 1 << 14
    rp_h, col_h = orc.hypercube_csr(m)
    fo, _ = orc.sobol_eval(sv, functor, gset, 0, m, b[:, 0], b[:, 1])
    sub = hp.star_csr(torch.from_numpy(rp_h).cuda(), torch.from_numpy(col_h).cuda(),
                      hp.eval_sobol(fid, gid, tab, 0, m)[0]).cpu().numpy()
    ref = orc.star_csr(rp_h, col_h, fo)
    if functor in DEVICE_BIT_EXACT:
        assert np.array_equal(sub, ref)
    else:
        # acceptance criterion: pool differences must be traced to near-ties
        from shgo_b200.ties import explain_pool_difference
        fd = hp.eval_sobol(fid, gid, tab, 0, m)[0].cpu().numpy()
        explained, report = explain_pool_difference(rp_h, col_h, fd, fo, sub, ref, max_ulps=8)
        assert explained, report[:5]


# ------------------------------------------------------------------------------------------
# local bounds of pool vertices (construct_lcb_simplicial, _shgo.py:1246-1270)
def _local_bounds_names():
    return [str(s) for s in load_golden("local_bounds")["names"]]


@pytest.mark.parametrize("name", _local_bounds_names())
def test_local_bounds_csr_golden(hp, name):
    """explicit graph: every vertex of every fixture, bit-equal to the reference's own method"""
    g = load_golden("local_bounds")
    x, edges, bounds = g[name + "_x"], g[name + "_edges"], g[name + "_bounds"]
    n = len(x)
    rp, col = orc.csr_from_edges(edges, n)
    xs = torch.from_numpy(np.ascontiguousarray(x.T)).cuda()
    pool = torch.arange(n, dtype=torch.int64).cuda()
    lo, hi = hp.local_bounds_csr(torch.from_numpy(rp).cuda(), torch.from_numpy(col).cuda(), xs, pool, bounds)
    assert np.array_equal(lo.cpu().numpy(), g[name + "_lo"])
    assert np.array_equal(hi.cpu().numpy(), g[name + "_hi"])
    # a sub-pool in arbitrary order, with repeats
    sub = np.random.default_rng(3).integers(0, n, 17)
    lo, hi = hp.local_bounds_csr(torch.from_numpy(rp).cuda(), torch.from_numpy(col).cuda(), xs,
                                 torch.from_numpy(sub).cuda(), bounds)
    assert np.array_equal(lo.cpu().numpy(), g[name + "_lo"][sub])
    assert np.array_equal(hi.cpu().numpy(), g[name + "_hi"][sub])


def test_local_bounds_of_an_isolated_vertex(hp):
    """shgo(..., n=1, sampling_method='simplicial'): one vertex, no edges, no column array at all"""
    rp = torch.zeros(2, dtype=torch.int64, device="cuda")
    col = torch.zeros(0, dtype=torch.int32, device="cuda")
    x = torch.tensor([[0.5], [2.0]], dtype=torch.float64, device="cuda")
    lo, hi = hp.local_bounds_csr(rp, col, x, torch.zeros(1, dtype=torch.int64, device="cuda"), [(-1, 6), (-1, 6)])
    assert lo.cpu().tolist() == [[-1.0, -1.0]] and hi.cpu().tolist() == [[6.0, 6.0]]


@pytest.mark.parametrize("name", [n for n in _local_bounds_names() if n.startswith(("lat2d", "lat4d"))])
def test_lattice_local_bounds_golden(hp, name):
    """whole generations: closed form on the coordinate table == the reference walking v.nn"""
    from shgo_b200 import lattice
    g = load_golden("local_bounds")
    x, bounds = g[name + "_x"], [tuple(b) for b in g[name + "_bounds"]]
    d, gen = len(bounds), int(name.rsplit("_g", 1)[1])
    tab = torch.from_numpy(lattice.coordinate_tables(bounds, gen)).cuda()
    nv, _ = orc.lattice_nvert(d, gen)
    assert nv == len(x)
    mine = hp.lattice_coords(d, gen, tab, 0, nv).cpu().numpy().T
    where = {tuple(c): i for i, c in enumerate(x)}
    order = np.array([where[tuple(c)] for c in mine])       # golden row of lattice id i
    lo, hi = hp.lattice_local_bounds(d, gen, tab, torch.arange(nv, dtype=torch.int64).cuda(), bounds)
    assert np.array_equal(lo.cpu().numpy(), g[name + "_lo"][order])
    assert np.array_equal(hi.cpu().numpy(), g[name + "_hi"][order])


def test_lattice_local_bounds_vs_explicit_walk_10d(hp):
    """10-D generation 1 (60 073 vertices, stars of up to 59 048): closed form == oracle walk"""
    d, gen = 10, 1
    bounds = [(-32.768, 32.768)] * d
    from shgo_b200 import lattice
    tab_h = lattice.coordinate_tables(bounds, gen)
    nv, _ = orc.lattice_nvert(d, gen)
    pool = np.random.default_rng(5).integers(0, nv, 300)
    pool[:3] = [0, nv - 1, nv // 2]
    x = orc.lattice_coords(bounds, gen)
    rp, col = orc.lattice_csr(d, gen)
    want_lo, want_hi = orc.local_bounds_csr(rp, col, x, pool, bounds)
    lo, hi = hp.lattice_local_bounds(d, gen, torch.from_numpy(tab_h).cuda(), torch.from_numpy(pool).cuda(), bounds)
    assert np.array_equal(lo.cpu().numpy(), want_lo) and np.array_equal(hi.cpu().numpy(), want_hi)
    assert hp.lattice_local_bounds(d, gen, torch.from_numpy(tab_h).cuda(),
                                   torch.zeros(0, dtype=torch.int64).cuda(), bounds)[0].shape == (0, d)


@pytest.mark.parametrize("nparts", [2, 3, 8])
def test_lattice_granule_cyclic_parts_emulated_on_one_gpu(hp, nparts):
    """Multi-GPU lattice path with every 'GPU' living on this one: each part evaluates its granules
    and pushes them into all copies of f (peer pointers = other local vectors); the union of the
    parts' flags must equal the single-launch result.  (Real multi-process path:
    tools/check_sharded_lattice.py under torchrun.)"""
    from shgo_b200 import lattice
    d, gen = 6, 2                      # 15 625 + 4 096 = 19 721 vertices: 5 granules, last one partial
    bounds = [(-32.768, 32.768)] * d
    tab = torch.from_numpy(lattice.coordinate_tables(bounds, gen)).cuda()
    nv, _ = orc.lattice_nvert(d, gen)
    fid, gid = _ids("ackley", None)
    f_ref, feas_ref = hp.lattice_eval(fid, gid, d, gen, tab, 0, nv)
    ref = hp.lattice_star(d, gen, f_ref)
    copies = [torch.full((nv,), float("nan"), dtype=torch.float64, device="cuda") for _ in range(nparts)]
    feas = [torch.zeros(nv, dtype=torch.uint8, device="cuda") for _ in range(nparts)]
    for p in range(nparts):
        others = [c.data_ptr() for q, c in enumerate(copies) if q != p]
        hp.lattice_eval_push(fid, gid, d, gen, tab, 0, nv, copies[p], feas[p], others, p, nparts, 4096)
    for c in copies:                   # every copy is complete and bit-identical
        assert torch.equal(c, f_ref)
    assert torch.equal(sum(feas), feas_ref)
    union = torch.zeros(nv, dtype=torch.uint8, device="cuda")
    owner = (torch.arange(nv, device="cuda") // 4096) % nparts
    for p in range(nparts):
        fl = hp.lattice_star(d, gen, copies[p], part=p, nparts=nparts, granule=4096)
        assert int(fl[owner != p].sum().item()) == 0
        union += fl
    assert torch.equal(union, ref)
    assert np.array_equal(ref.cpu().numpy(), orc.lattice_star(d, gen, f_ref.cpu().numpy()))


# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d,n", [(1, 1000), (2, 70001), (3, 4096), (6, 20000), (8, 5000), (40, 3000)])
def test_morton_order_is_a_stable_sort_by_the_z_order_key(hp, d, n):
    rng = np.random.default_rng(d)
    bounds = np.stack([rng.uniform(-5, 0, d), rng.uniform(1, 9, d)], 1)
    pts = rng.uniform(bounds[:, 0], bounds[:, 1], size=(n, d))
    pts[::97] = pts[1]                                # coincident points: ties keep index order
    pts[5] = bounds[:, 1]                             # a point on the upper faces
    x = torch.from_numpy(pts).cuda().t().contiguous()
    order, newid = hp.morton_order(x, bounds)
    order, newid = order.cpu().numpy(), newid.cpu().numpy()
    key = orc.morton_keys(pts, bounds)
    assert np.array_equal(order, np.argsort(key, kind="stable"))
    assert np.array_equal(newid[order], np.arange(n))


def test_star_test_in_morton_numbering_gives_the_same_pool(hp):
    """real qhull meshes (2-D and 3-D Sobol points): CSR in the mesh's own numbering vs in Z-order,
    pools identical after mapping back, and identical to the oracle's"""
    from scipy.spatial import Delaunay
    for d, n in ((2, 1 << 14), (3, 1 << 12)):
        bounds = [(-5.12, 5.12)] * d
        tab = hp.SobolTables(bounds)
        pts = hp.sobol_points(tab, 0, n, "aos").cpu().numpy()
        simp = Delaunay(pts).simplices.astype(np.int32)
        simp_d = torch.from_numpy(simp).cuda()
        f, _ = hp.eval_sobol(orc.FUNCTOR_IDS["rastrigin"], 0, tab, 0, n)
        rp, col = hp.csr_from_simplices(simp_d, n)
        flags = hp.star_csr(rp, col, f)
        x = hp.sobol_points(tab, 0, n, "soa")
        order, newid = hp.morton_order(x, bounds)
        s3 = hp.relabel3(simp_d, n, newid)
        assert np.array_equal(s3.cpu().numpy(), newid.cpu().numpy()[simp[:, :3]])
        rp_m, col_m = hp.csr_from_simplices(s3, n)
        assert int(rp_m[-1]) == int(rp[-1])
        flags_m = hp.star_csr(rp_m, col_m, hp.gather(f, order))
        back = hp.scatter_u8(flags_m, order, torch.empty(n, dtype=torch.uint8, device="cuda"))
        assert torch.equal(back, flags)
        rp_o, col_o = orc.vf_to_vv(simp, n)
        assert np.array_equal(back.cpu().numpy(), orc.star_csr(rp_o, col_o, f.cpu().numpy()))
        # locality: mean |row - neighbour| drops by an order of magnitude
        def spread(rpt, ct):
            rows = np.repeat(np.arange(n), np.diff(rpt.cpu().numpy()))
            return np.abs(rows - ct.cpu().numpy()).mean()
        assert spread(rp_m, col_m) * 8 < spread(rp, col)


def test_csr_rows_longer_than_a_warp_are_sorted(hp):
    """hubs: rows of 33 ... 50 000 entries go through the per-row radix sort"""
    rng = np.random.default_rng(3)
    n = 60000
    hubs = [0, 17, 40000]
    sizes = [33, 700, 50000]
    tri = []
    for h, m in zip(hubs, sizes):
        others = rng.choice(np.setdiff1d(np.arange(n), hubs), m, replace=False)
        tri.append(np.stack([np.full(m, h), others, np.roll(others, 1)], 1))
    tri.append(rng.integers(0, n, size=(30000, 3)))
    simp = np.concatenate(tri).astype(np.int32)
    rp, col = hp.csr_from_simplices(torch.from_numpy(simp).cuda(), n)
    rp_o, col_o = orc.vf_to_vv(simp, n)
    assert np.array_equal(rp.cpu().numpy(), rp_o) and np.array_equal(col.cpu().numpy(), col_o)


def test_argmin_breaks_ties_by_rank(hp):
    f = torch.tensor([3.0, 1.0, float("inf"), 1.0, 2.0, 1.0, float("nan")], dtype=torch.float64, device="cuda")
    assert hp.argmin(f) == (1, 1.0)
    rank = torch.tensor([5, 9, 0, 7, 1, 2, 3], dtype=torch.int64, device="cuda")
    assert hp.argmin(f, None, rank) == (5, 1.0)
    present = torch.tensor([1, 1, 1, 1, 1, 0, 1], dtype=torch.uint8, device="cuda")
    assert hp.argmin(f, present, rank) == (3, 1.0)
    big = torch.full((1 << 20,), 2.5, dtype=torch.float64, device="cuda")
    r = torch.arange(1 << 20, 0, -1, dtype=torch.int64, device="cuda")
    assert hp.argmin(big, None, r)[0] == (1 << 20) - 1 and hp.argmin(big)[0] == 0
    assert hp.argmin(torch.full((10,), float("inf"), dtype=torch.float64, device="cuda")) == (-1, float("inf"))


def test_sampling_subspace_on_the_device(hp):
    """row a4': C = C[g(C.T) >= 0.0] for every g (reference _shgo.py:1452-1463); NaN drops the point"""
    from shgo_b200 import functors as F
    n, d = 5000, 4
    tab = hp.SobolTables([(0.0, 1.0)] * d)
    x = hp.sobol_points(tab, 0, n, "soa")
    C = x.t().contiguous().cpu().numpy()
    gs = (F.cattle_feed_g1, F.cattle_feed_g2)
    g = torch.stack([torch.from_numpy(np.asarray(gi(C.T))).cuda() for gi in gs])
    g[0, 7] = float("nan")
    g[1, 11] = -0.0                                     # -0.0 >= 0.0 is True
    idx, kept = hp.sampling_subspace(x, g)
    gh = g.cpu().numpy()
    ref_keep = np.all(gh >= 0.0, axis=0)
    assert not ref_keep[7]
    assert np.array_equal(idx.cpu().numpy(), np.flatnonzero(ref_keep))
    assert np.array_equal(kept.cpu().numpy(), C[ref_keep])
    idx2, C2 = orc.sampling_subspace(C, gs)             # the reference's own loop
    ok = np.ones(n, bool); ok[7] = False
    assert np.array_equal(np.intersect1d(idx2, np.flatnonzero(ok)), np.intersect1d(idx.cpu().numpy(), np.flatnonzero(ok)))
    e_idx, e_pts = hp.sampling_subspace(x[:, :0].contiguous(), None)
    assert e_idx.numel() == 0 and e_pts.shape == (0, d)


@pytest.mark.parametrize("m,ties", [(1, False), (40, False), (5000, True), (300000, True)])
def test_pool_rank_on_the_device(hp, m, ties):
    """row f3: LMC exclusion by coordinates + f-sorted pool + distance-sorted pool"""
    rng = np.random.default_rng(m)
    n, d = max(4 * m, 64), 3
    x = rng.uniform(-5, 5, (n, d))
    f = rng.standard_normal(n)
    if ties:
        f = np.round(f, 2)                              # many exactly equal values
    f[rng.integers(0, n, 5)] = np.inf
    f[0] = -0.0
    pool = np.sort(rng.choice(n, m, replace=False))
    xl = [x[pool[len(pool) // 2]].copy(), x[pool[0]].copy(), np.array([9.0, 9.0, 9.0])]
    xs = torch.from_numpy(x).cuda().t().contiguous()
    fd = torch.from_numpy(f).cuda()
    pd = torch.from_numpy(pool).cuda()
    got = hp.pool_rank(pd, fd, xs, xl).cpu().numpy()
    assert np.array_equal(got, orc.pool_rank(pool, f, x, xl))
    assert np.array_equal(hp.pool_rank(pd, fd).cpu().numpy(), orc.pool_rank(pool, f))
    xref = x[pool[-1]]
    got = hp.pool_rank(pd, fd, xs, xl[:1], by="distance", xref=xref).cpu().numpy()
    assert np.array_equal(got, orc.pool_rank(pool, f, x, xl[:1], by="distance", xref=xref))
