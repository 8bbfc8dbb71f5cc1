"""HOST logic of DeviceComplex and the shgo() wrapper, on CPU, with the oracle standing in for the
kernels (tests/fake_hotpath.py).  The same scenarios run against the real kernels in
tests/test_gpu_api.py.  Comparator: the driver with its own Python Complex, same process."""
import numpy as np
import pytest

import fake_hotpath
from conftest import load_golden


@pytest.fixture()
def sb(monkeypatch):
    fake_hotpath.install(monkeypatch)
    import shgo_b200
    from shgo_b200 import api
    yield shgo_b200
    api.uninstall()


def _reference_run(fn, *a, **k):
    from shgo_b200 import api
    m = api.driver_module()
    assert getattr(m, "_shgo_b200_saved_complex", None) is None
    return m.shgo(fn, *a, **k)


def test_readme_example(sb):
    from scipy.optimize import rosen
    g = load_golden("end_to_end")
    res = sb.shgo(rosen, [(0, 2)] * 5)
    assert res.success and res.nfev == int(g["rosen5_nfev"])
    np.testing.assert_allclose(res.x, g["rosen5_x"], atol=1e-8)


def test_cattle_feed(sb):
    from shgo_b200 import functors as F
    g = load_golden("end_to_end")
    cons = ({"type": "ineq", "fun": F.cattle_feed_g1}, {"type": "ineq", "fun": F.cattle_feed_g2},
            {"type": "eq", "fun": F.cattle_feed_h1})
    res = sb.shgo(F.cattle_feed, [(0, 1.0)] * 4, n=150, constraints=cons)
    assert res.success and res.nfev == int(g["cattle_nfev"])
    assert res.fun == pytest.approx(float(g["cattle_fun"]), rel=1e-9)


def test_eggholder_sobol(sb):
    from shgo_b200 import functors as F
    g = load_golden("end_to_end")
    res = sb.shgo(F.eggholder, [(-512, 512)] * 2, n=64, sampling_method="sobol")
    assert len(res.funl) == len(g["egg_funl"]) and res.nfev == int(g["egg_nfev"])
    np.testing.assert_allclose(res.funl, g["egg_funl"], rtol=1e-8)


KW = [dict(n=None, iters=3, sampling_method="simplicial"),
      dict(n=30, iters=3, sampling_method="simplicial"),
      dict(n=100, iters=2, sampling_method="simplicial"),
      dict(n=128, iters=1, sampling_method="sobol"),
      dict(n=64, iters=3, sampling_method="sobol")]


@pytest.mark.parametrize("kw", KW)
@pytest.mark.parametrize("problem", ["rastrigin2", "ackley3", "sphere_cons", "python_callable"])
def test_same_result_as_driver_with_python_complex(sb, kw, problem):
    from shgo_b200 import functors as F
    if problem == "rastrigin2":
        fn, bounds, cons = F.rastrigin, [(-5.12, 5.12)] * 2, None
    elif problem == "ackley3":
        fn, bounds, cons = F.ackley, [(-32.768, 32.768)] * 3, None
    elif problem == "sphere_cons":
        fn, bounds, cons = F.sphere, [(-1, 6)] * 2, ({"type": "ineq", "fun": F.sum_le_6},)
    else:
        def fn(x):
            return (x[0] - 0.3) ** 2 + np.sin(3 * x[1]) + 0.1 * x[0] * x[1]
        bounds, cons = [(-2, 2), (-3, 1)], None
    ref = _reference_run(fn, bounds, constraints=cons, **kw)
    res = sb.shgo(fn, bounds, constraints=cons, **kw)
    assert res.success == ref.success
    assert res.nfev == ref.nfev and res.nit == ref.nit
    assert len(res.funl) == len(ref.funl)
    np.testing.assert_allclose(res.funl, ref.funl, rtol=1e-7, atol=1e-9)
    # minima with exactly equal f (symmetric problems) may come out in a different order: the
    # reference breaks such ties by the insertion order of its vertex cache (documented deviation)
    key = lambda a: np.lexsort(np.round(np.asarray(a), 5).T[::-1])
    np.testing.assert_allclose(np.asarray(res.xl)[key(res.xl)], np.asarray(ref.xl)[key(ref.xl)],
                               rtol=1e-6, atol=1e-6)


def test_pool_objects_and_lazy_cache(sb):
    from shgo_b200 import functors as F
    s = sb.SHGO(F.rastrigin, [(-5.12, 5.12)] * 3, n=None, iters=3)
    s.iterate()
    s.iterate()
    s.iterate_complex()
    s.minimizers()
    V = s.HC.V
    assert s.HC.mode == "lattice" and V.size() == 189 and V.nfev == 189
    assert len(V.cache) < 60                         # only pool + arg-min materialised, not 189
    for v in s.minimizer_pool:
        assert v.minimiser() and all(v.f < w.f for w in v.nn)
        assert V[v.x] is v and v.x_a.shape == (3,)


def test_infeasible_and_flat_problems(sb):
    def table(x):
        return 0.0 if x[0] > 3.0 else 3.0
    res = sb.shgo(table, [(-10, 10)] * 2, n=64, sampling_method="sobol")
    ref = _reference_run(table, [(-10, 10)] * 2, n=64, sampling_method="sobol")
    assert res.success == ref.success and res.nfev == ref.nfev

    def g_bad(x):
        return -1.0 - x[0] ** 2
    cons = ({"type": "ineq", "fun": g_bad},)
    res = sb.shgo(lambda x: x[0] ** 2 + x[1] ** 2, [(-1, 1)] * 2, constraints=cons, n=32,
                  sampling_method="sobol")
    assert not res.success and res.nfev == 0


@pytest.mark.parametrize("bounds", [[(-3.3, 7.1)] * 3, [(0.1, 0.7)] * 2, [(0.3, 1.9), (-3.3, 7.1)]])
def test_lattice_artefact_bounds_follow_the_reference(sb, bounds):
    """Bounds whose fp64 midpoints depend on the split direction: the reference builds near-duplicate
    vertices there, the closed-form lattice does not apply, and the complex replays the reference's
    own generator instead -- same vertex count, nfev and minima as the reference."""
    from shgo_b200 import functors as F
    kw = dict(n=None, iters=3, sampling_method="simplicial")
    ref = _reference_run(F.sphere, bounds, **kw)
    res = sb.shgo(F.sphere, bounds, **kw)
    assert res.nfev == ref.nfev and res.nit == ref.nit and res.success == ref.success
    np.testing.assert_allclose(np.sort(res.funl), np.sort(ref.funl), rtol=1e-9, atol=1e-12)
    from shgo_b200 import api
    s = sb.SHGO(F.sphere, bounds, **kw)
    try:
        s.iterate_complex()
        s.iterate_complex()
        s.iterate_complex()
    finally:
        api.uninstall()
    m = api.driver_module()
    r = m.SHGO(F.sphere, bounds, **kw)
    r.iterate_complex()
    r.iterate_complex()
    r.iterate_complex()
    assert s.HC.mode == "hostgraph" and s.HC.V.size() == r.HC.V.size() and s.HC.V.nfev == r.HC.V.nfev


@pytest.mark.parametrize("kw", [dict(n=None, iters=3), dict(n=30, iters=3), dict(n=70, iters=2)])
def test_local_bounds_through_the_driver_seam(sb, kw):
    """SHGO.construct_lcb_simplicial served by DeviceComplex.local_bounds --
    identical boxes, and no neighbour object is created for them."""
    from shgo_b200 import api, functors as F
    s = sb.SHGO(F.styblinski_tang, [(-5.0, 4.0)] * 3, sampling_method="simplicial", **kw)
    s.iterate_complex()
    s.iterate_complex()
    s.minimizers()
    m = api.driver_module()
    reference_method = getattr(m.SHGO, "_shgo_b200_saved_lcb", m.SHGO.construct_lcb_simplicial)
    assert len(s.minimizer_pool) > 0
    boxes = [s.construct_lcb_simplicial(v) for v in s.minimizer_pool]
    assert len(s.HC.V._unlisted) == 0
    for v, got in zip(s.minimizer_pool, boxes):
        assert got == reference_method(s, v)          # walks v.nn: materialises the star
        assert all(lo <= x <= hi for (lo, hi), x in zip(got, v.x))
    assert len(s.HC.V._unlisted) > 0


def test_coincident_points_merge_like_the_reference(sb):
    """Vertices are keyed by coordinates in the reference (_vertex.py:304-317): points that appear
    twice in a mesh are ONE vertex.  Same (points, simplices) into DeviceComplex.vf_to_vv and into
    the driver's Python Complex: same vertex set, same minimisers, same nfev."""
    from scipy.spatial import Delaunay
    from shgo_b200 import api, functors as F
    from shgo_b200.complex import DeviceComplex
    rng = np.random.default_rng(11)
    base = rng.uniform(-5.12, 5.12, size=(200, 2))
    tri = Delaunay(base).simplices.astype(np.int32)
    # 40 points get a second copy (new index, same coordinates); half of the simplices that used
    # the original now reference the copy
    dup = rng.choice(200, 40, replace=False)
    pts = np.concatenate([base, base[dup]])
    copy_of = {int(j): 200 + k for k, j in enumerate(dup)}
    simp = tri.copy()
    for s in range(0, len(simp), 2):
        simp[s] = [copy_of.get(int(v), int(v)) for v in simp[s]]
    bounds = [(-5.12, 5.12)] * 2
    m = api.driver_module()
    RefComplex = getattr(m, "_shgo_b200_saved_complex", None) or m.Complex
    ref = RefComplex(2, domain=bounds, sfield=F.rastrigin)
    ref.vf_to_vv(pts, simp)
    ref.V.process_pools()
    ref_min = sorted(v.x for v in ref.V.cache.values() if v.minimiser())
    dev = DeviceComplex(2, domain=bounds, sfield=F.rastrigin)
    dev.vf_to_vv(pts, simp)
    dev.V.process_pools()
    assert dev.V.size() == len(ref.V.cache) == 200
    assert dev.V.nfev == ref.V.nfev
    got_min = sorted(v.x for v in dev.V.cache.values() if v.minimiser())
    assert got_min == ref_min and len(ref_min) > 3
    for v in dev.V.cache.values():            # neighbour sets, by coordinates
        if v.minimiser():
            assert {w.x for w in v.nn} == {w.x for w in ref.V[v.x].nn}


REGRESSIONS = [
    # 1-D meshes: every iteration hands over only its fresh points (driver: Tri = (C, edges))
    ("rastrigin", [(-2.1, 4.4)], dict(sampling_method="sobol", n=94, iters=2)),
    ("styblinski_tang", [(-2.0, 6.2)], dict(sampling_method="halton", n=86, iters=2)),
    # the reference's star() self-loop changes the edges later partial refinements create
    ("ackley", [(-6.0, -2.73), (-4.0, -2.13)], dict(sampling_method="simplicial", n=40, iters=3)),
    # the literal default call with several iterations: generation 0 in closed form, then partial
    ("rastrigin", [(-5.12, 4.0)] * 2, dict(sampling_method="simplicial", n=5, iters=4)),
    ("styblinski_tang", [(-5.0, 4.0)] * 3, dict(sampling_method="simplicial", n=9, iters=3)),
    ("ackley", [(-20.0, 32.768)] * 4, dict(sampling_method="simplicial", iters=2)),
    # options['symmetry']: the generator must get the driver's own ndarray of bounds
    ("sphere", [(-4.04, 4.46), (-1.59, 5.91), (-0.93, 2.57), (-4.07, 3.43)],
     dict(sampling_method="simplicial", n=6, iters=2, options={"symmetry": True})),
]


@pytest.mark.parametrize("fname,bounds,kw", REGRESSIONS)
def test_cases_found_by_the_differential_fuzz(sb, fname, bounds, kw):
    from shgo_b200 import functors as F
    fn = getattr(F, fname)
    ref = _reference_run(fn, bounds, **kw)
    res = sb.shgo(fn, bounds, **kw)
    assert (res.success, res.nfev, res.nit, len(res.funl)) == (ref.success, ref.nfev, ref.nit, len(ref.funl))
    np.testing.assert_allclose(np.sort(res.funl), np.sort(ref.funl), rtol=1e-6, atol=1e-8)


def test_shgo_object_does_not_leave_the_driver_patched(sb):
    """shgo_b200.SHGO(...) binds the device complex for the construction only: a later plain driver
    call in the same process gets the driver's own Python Complex again."""
    from shgo_b200 import api, functors as F
    from shgo_b200.complex import DeviceComplex
    m = api.driver_module()
    plain = m.Complex
    assert plain is not DeviceComplex
    s = sb.SHGO(F.sphere, [(-1.0, 2.0)] * 2, n=None, iters=1)
    assert isinstance(s.HC, DeviceComplex)
    assert m.Complex is plain and not hasattr(m.SHGO, "_shgo_b200_saved_lcb")
    r = m.SHGO(F.sphere, [(-1.0, 2.0)] * 2, n=None, iters=1)
    assert not isinstance(r.HC, DeviceComplex)
    s.iterate_all()
    r.iterate_all()
    assert s.res.nfev == r.res.nfev and np.allclose(s.res.x, r.res.x)


def test_mesh_classified_in_morton_numbering_gives_the_same_result(sb, monkeypatch):
    """Large meshes are also kept in a Z-order numbering for the star test (complex.MORTON_MIN_VERTICES):
    the driver must not see any difference -- same result object with the threshold forced to 0."""
    from shgo_b200 import complex as cx, functors as F
    kw = dict(n=128, iters=2, sampling_method="sobol")
    bounds = [(-5.12, 5.12)] * 3
    plain = sb.shgo(F.rastrigin, bounds, **kw)
    monkeypatch.setattr(cx, "MORTON_MIN_VERTICES", 0)
    used = []
    real = cx.hp.morton_order
    monkeypatch.setattr(cx.hp, "morton_order", lambda *a, **k: used.append(1) or real(*a, **k))
    morton = sb.shgo(F.rastrigin, bounds, **kw)
    assert used, "the Z-order numbering was not exercised"
    assert (plain.nfev, plain.nit, plain.nlfev) == (morton.nfev, morton.nit, morton.nlfev)
    assert np.array_equal(plain.xl, morton.xl) and np.array_equal(plain.funl, morton.funl)


def test_large_pools_are_ranked_on_the_device_with_the_same_outcome(sb, monkeypatch):
    """SHGO.minimizers' LMC exclusion + f-ordering served by DeviceComplex.ranked_pool
    (threshold forced to 0 here): same result as the driver's own loop + np.argsort."""
    from shgo_b200 import api, functors as F
    bounds = [(-5.12, 4.7), (-4.9, 5.12)]                 # no symmetry: no exactly equal f values
    kw = dict(n=None, iters=3, sampling_method="simplicial")
    plain = sb.shgo(F.rastrigin, bounds, **kw)
    monkeypatch.setattr(api, "DEVICE_POOL_RANK_MIN", 0)
    used = []
    real = api._cx.DeviceComplex.ranked_pool
    monkeypatch.setattr(api._cx.DeviceComplex, "ranked_pool",
                        lambda self, *a, **k: used.append(1) or real(self, *a, **k))
    dev = sb.shgo(F.rastrigin, bounds, **kw)
    assert used
    assert (plain.nfev, plain.nit, plain.nlfev, plain.success) == (dev.nfev, dev.nit, dev.nlfev, dev.success)
    assert np.array_equal(plain.xl, dev.xl) and np.array_equal(plain.funl, dev.funl)
    kw = dict(n=256, iters=2, sampling_method="sobol")
    plain2 = _reference_run(F.styblinski_tang, [(-5.0, 4.0)] * 3, **kw)
    dev2 = sb.shgo(F.styblinski_tang, [(-5.0, 4.0)] * 3, **kw)
    assert np.array_equal(plain2.xl, dev2.xl) and plain2.nfev == dev2.nfev


def test_sampling_subspace_served_by_the_device(sb, monkeypatch):
    """row a4': options['infty_constraints'] = False makes the driver drop infeasible samples before
    triangulating (_shgo.py:1452-1463); with a registered constraint set that filter runs on the device"""
    from shgo_b200 import api, functors as F
    cons = ({"type": "ineq", "fun": F.cattle_feed_g1}, {"type": "ineq", "fun": F.cattle_feed_g2})
    kw = dict(n=128, iters=1, sampling_method="sobol", constraints=cons, options={"infty_constraints": False})
    ref = _reference_run(F.cattle_feed, [(0.0, 1.0)] * 4, **kw)
    used = []
    real = api._cx.device_sampling_subspace
    monkeypatch.setattr(api._cx, "device_sampling_subspace", lambda *a, **k: used.append(1) or real(*a, **k))
    res = sb.shgo(F.cattle_feed, [(0.0, 1.0)] * 4, **kw)
    assert used
    assert (res.nfev, res.nit, res.success) == (ref.nfev, ref.nit, ref.success)
    np.testing.assert_allclose(res.x, ref.x, rtol=1e-9, atol=1e-12)


def test_lattice_limits_fall_back_to_the_generator(sb):
    """the closed-form lattice kernels take d <= 12 and < 2^32 vertices; beyond that the
    complex must not enter lattice mode (the caller then replays the generator, like for artefact bounds)"""
    from shgo_b200 import functors as F
    from shgo_b200.complex import DeviceComplex
    c13 = DeviceComplex(13, domain=[(-1.0, 1.0)] * 13, sfield=F.sphere)
    assert c13._enter_lattice(0) is False and c13.mode is None
    c12 = DeviceComplex(12, domain=[(-1.0, 1.0)] * 12, sfield=F.sphere)
    assert c12._enter_lattice(3) is False          # (2^3+1)^12 + 2^36 vertices > 2^32
    assert c12._enter_lattice(0) is True and c12.mode == "lattice"
