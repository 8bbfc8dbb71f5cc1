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
