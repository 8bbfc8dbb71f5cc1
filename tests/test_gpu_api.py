"""The reference's public API on the device complex: `shgo_b200.shgo(...)` against results recorded
from the live reference (tests/golden/end_to_end.npz) and against the driver with its own Python
Complex run side by side (same driver package, same process).  B200 only."""
import numpy as np
import pytest

from conftest import load_golden

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sb():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import shgo_b200
    return shgo_b200


def _reference_run(fn, *a, **k):
    """Same call through the driver's own pure-Python Complex (the oracle side of the API test)."""
    from shgo_b200 import api
    m = api.driver_module()
    assert getattr(m, "_shgo_b200_saved_complex", None) is None, "device complex still installed"
    return m.shgo(fn, *a, **k)


def test_readme_example_rosenbrock(sb):
    """config 1: shgo(rosen, [(0,2)]*5) -> x = 1, fun = 2.92e-18, nfev = 50 (_shgo.py:307-311)"""
    from scipy.optimize import rosen
    g = load_golden("end_to_end")
    res = sb.shgo(rosen, [(0, 2)] * 5)
    assert res.success
    np.testing.assert_allclose(res.x, g["rosen5_x"], atol=1e-8)
    assert res.fun == pytest.approx(float(g["rosen5_fun"]), abs=1e-12)
    assert res.nfev == int(g["rosen5_nfev"])


def test_cattle_feed_constrained(sb):
    from shgo_b200 import functors as F
    g = load_golden("end_to_end")
    cons = ({"type": "ineq", "fun": F.cattle_feed_g1}, {"type": "ineq", "fun": F.cattle_feed_g2},
            {"type": "eq", "fun": F.cattle_feed_h1})
    res = sb.shgo(F.cattle_feed, [(0, 1.0)] * 4, n=150, constraints=cons)
    assert res.success
    assert res.fun == pytest.approx(float(g["cattle_fun"]), rel=1e-9)
    np.testing.assert_allclose(res.x, g["cattle_x"], atol=1e-6)
    assert res.nfev == int(g["cattle_nfev"])


def test_eggholder_sobol(sb):
    from shgo_b200 import functors as F
    g = load_golden("end_to_end")
    res = sb.shgo(F.eggholder, [(-512, 512)] * 2, n=64, sampling_method="sobol")
    assert res.success
    assert len(res.funl) == len(g["egg_funl"])
    np.testing.assert_allclose(res.funl, g["egg_funl"], rtol=1e-8)
    np.testing.assert_allclose(res.xl, g["egg_xl"], rtol=1e-6, atol=1e-6)
    assert res.nfev == int(g["egg_nfev"])


@pytest.mark.parametrize("kw", [
    dict(n=None, iters=3, sampling_method="simplicial"),          # whole generations: lattice kernels
    dict(n=30, iters=3, sampling_method="simplicial"),            # partial generations: host graph
    dict(n=128, iters=1, sampling_method="sobol"),
    dict(n=64, iters=3, sampling_method="sobol"),                 # engine continuation + re-triangulation
])
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
        def fn(x):                          # arbitrary user code: evaluated on the host, per vertex
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


def test_torch_vectorised_callable(sb):
    from shgo_b200.complex import torch_vectorized

    @torch_vectorized
    def f(x):
        return (x[0] - 1.0) ** 2 + (x[1] + 0.5) ** 2 + torch.cos(3.0 * x[0])

    def f_np(x):
        return (x[0] - 1.0) ** 2 + (x[1] + 0.5) ** 2 + np.cos(3.0 * x[0])

    s = sb.SHGO(f, [(-3, 3)] * 2, n=512, iters=1, sampling_method="sobol")
    s.iterate_complex()
    s.minimizers()
    pool = sorted(tuple(v.x) for v in s.minimizer_pool)
    m = __import__("shgo_b200.api", fromlist=["api"]).uninstall()
    r = m.SHGO(f_np, [(-3, 3)] * 2, n=512, iters=1, sampling_method="sobol")
    r.iterate_complex()
    r.minimizers()
    assert pool == sorted(tuple(v.x) for v in r.minimizer_pool)
    assert s.HC.V.nfev == r.HC.V.nfev


def test_infeasible_and_flat_problems(sb):
    """reference tests/test__shgo.py:932-998: no pool / empty feasible set end in a clean failure"""
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


@pytest.mark.parametrize("kw", [dict(n=None, iters=3), dict(n=30, iters=3), dict(n=70, iters=2)])
def test_local_bounds_through_the_driver_seam(sb, kw):
    """SHGO.construct_lcb_simplicial served by the device kernels: identical
    boxes (lattice closed form for whole generations, CSR walk for partial ones), no neighbour
    objects created."""
    from shgo_b200 import api, functors as F
    s = sb.SHGO(F.styblinski_tang, [(-5.0, 4.0)] * 3, sampling_method="simplicial", **kw)
    try:
        s.iterate_complex()
        s.iterate_complex()
        s.minimizers()
        reference_method = getattr(api.driver_module().SHGO, "_shgo_b200_saved_lcb", api.driver_module().SHGO.construct_lcb_simplicial)
        assert len(s.minimizer_pool) > 0
        boxes = [s.construct_lcb_simplicial(v) for v in s.minimizer_pool]
        assert len(s.HC.V._unlisted) == 0
        for v, got in zip(s.minimizer_pool, boxes):
            assert got == reference_method(s, v)
    finally:
        api.uninstall()



def test_large_mesh_goes_through_the_z_order_numbering_and_device_pool_ranking(sb, monkeypatch):
    """A mesh of >= 2^15 vertices is classified in Z-order numbering (complex.MORTON_MIN_VERTICES) and a pool of
    >= DEVICE_POOL_RANK_MIN vertices is filtered / ranked on the device: the driver must see the same complex as
    with both switched off -- same vertex count, nfev, pool (coordinates, in order) and lowest vertex."""
    from shgo_b200 import api, complex as cx, functors as F
    bounds = [(-5.12, 4.7), (-4.9, 5.12)]
    kw = dict(n=1 << 15, iters=1, sampling_method="sobol")

    def run(morton_min, rank_min):
        monkeypatch.setattr(cx, "MORTON_MIN_VERTICES", morton_min)
        monkeypatch.setattr(api, "DEVICE_POOL_RANK_MIN", rank_min)
        s = sb.SHGO(F.rastrigin, bounds, **kw)
        s.iterate_complex()
        s.minimizers()
        return s

    a = run(1 << 40, 1 << 40)            # both off
    b = run(1 << 15, 16)                 # both on
    assert a.HC._morton is None and b.HC._morton is not None
    assert a.HC.V.size() == b.HC.V.size() and a.HC.V.nfev == b.HC.V.nfev
    assert np.array_equal(a.HC.pool_indices, b.HC.pool_indices) and len(a.HC.pool_indices) > 64
    assert a.HC.argmin_index == b.HC.argmin_index
    assert np.array_equal(a.X_min, b.X_min)
    assert np.array_equal(a.minimizer_pool_F, b.minimizer_pool_F)
    assert [v.x for v in a.minimizer_pool] == [v.x for v in b.minimizer_pool]
