"""Mesh edge cases through the real kernels (B200 only); the same scenarios run against the
oracle-backed stand-in in tests/test_host_complex.py."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sb():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import shgo_b200
    return shgo_b200


def _reference_run(fn, *a, **k):
    from shgo_b200 import api
    m = api.driver_module()
    assert getattr(m, "_shgo_b200_saved_complex", None) is None, "device complex still installed"
    return m.shgo(fn, *a, **k)


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
