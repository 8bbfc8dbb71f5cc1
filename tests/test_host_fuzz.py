"""Randomised differential test of the host logic: `shgo_b200.shgo` (DeviceComplex, kernels replaced
by the oracle-backed stand-in) against the driver with its own Python Complex on random problems
-- dimensions 1-4, uneven bounds, every sampling method, partial and whole generations, several
iterations, inequality and equality constraints, objective arguments, driver options (local_iter,
minimize_every_iter, maxfev / maxev / maxiter, infty_constraints, symmetry), local minimisers.
Found in round 1: 1-D meshes hand over only the fresh
points of an iteration; the reference's star() self-loops change which edges later partial
refinements create; with `symmetry` the reference's triangulate() behaves differently for an ndarray
of bounds than for a list of pairs.

`local_iter` starts its local searches from X_min[0], the pool vertex the reference's cache created
first; inside one whole generation that order has no closed form, so `shgo_b200.shgo` / `SHGO` give
such calls their graph from the reference's generator."""
import numpy as np
import pytest

import fake_hotpath


def _pyfun(x):
    return float(np.sum((x - 0.3) ** 2) + np.sin(3 * x[0]))


def _pyargs(x, a, b):
    return float(np.sum((x - a) ** 2) * b)


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4, 5])
def test_random_problems_match_the_driver_with_its_python_complex(monkeypatch, seed):
    _results_campaign(monkeypatch, seed, 25)


def test_random_problems_with_the_round_2_device_paths_forced(monkeypatch):
    """The same two campaigns with every mesh ALSO classified in Z-order numbering and every pool filtered /
    ranked by the device path (thresholds forced to 0): the driver must not see a difference.  (Run offline
    over ~700 problems against scipy's driver and the reference's own: no mismatch.)"""
    from shgo_b200 import api, complex as cx
    monkeypatch.setattr(cx, "MORTON_MIN_VERTICES", 0)
    monkeypatch.setattr(api, "DEVICE_POOL_RANK_MIN", 0)
    _results_campaign(monkeypatch, 41, 12)
    _iterations_campaign(monkeypatch, 41, 8)


def _results_campaign(monkeypatch, seed, trials):
    fake_hotpath.install(monkeypatch)
    import shgo_b200
    from shgo_b200 import api, functors as F
    m = api.driver_module()
    funs = [F.rastrigin, F.ackley, F.styblinski_tang, F.sphere, _pyfun, F.rosenbrock, _pyargs]
    rng = np.random.default_rng(1000 + seed)
    for trial in range(trials):
        d = int(rng.integers(1, 5))
        fn = funs[int(rng.integers(0, len(funs)))]
        if fn is F.rosenbrock and d < 2:
            d = 2
        lo = np.round(rng.uniform(-6, 0, d), int(rng.integers(0, 3)))
        hi = lo + np.round(rng.uniform(0.5, 8, d), int(rng.integers(0, 3))) + 0.5
        bounds = list(zip(lo.tolist(), hi.tolist()))
        method = ["simplicial", "sobol", "halton"][int(rng.integers(0, 3))]
        kw = dict(sampling_method=method)
        if fn is _pyargs:
            kw["args"] = (float(rng.uniform(-1, 1)), 2.0)
        if method == "simplicial":
            kw["n"] = [None, int(rng.integers(1, 60)), 2 ** d + 1][int(rng.integers(0, 3))]
            kw["iters"] = int(rng.integers(1, 4)) if d <= 3 else int(rng.integers(1, 3))
        else:
            kw["n"] = int(rng.integers(4, 100))
            kw["iters"] = int(rng.integers(1, 3))
        opts = {}
        if rng.random() < 0.3:
            opts["minimize_every_iter"] = bool(rng.integers(0, 2))
        if rng.random() < 0.3:
            opts["local_iter"] = int(rng.integers(1, 4))
        if rng.random() < 0.2:
            opts["maxfev"] = int(rng.integers(5, 200))
        if rng.random() < 0.15:
            opts["maxev"] = int(rng.integers(5, 200))
        if rng.random() < 0.15:
            opts["maxiter"] = int(rng.integers(1, 4))
        if rng.random() < 0.15:
            opts["infty_constraints"] = False
        if rng.random() < 0.1 and method == "simplicial":
            opts["symmetry"] = True
        if opts:
            kw["options"] = opts
        if rng.random() < 0.15:
            kw["minimizer_kwargs"] = {"method": ["SLSQP", "COBYLA", "L-BFGS-B"][int(rng.integers(0, 3))]}
        cons = None
        r = rng.random()
        if r < 0.25 and d >= 2:
            c = float(rng.uniform(-2, 6))
            cons = ({"type": "ineq", "fun": lambda x, *a, c=c: c - x[0] - x[1]},)
        elif r < 0.35 and d >= 2:
            c = float(rng.uniform(-2, 6))
            cons = ({"type": "ineq", "fun": lambda x, *a, c=c: c - x[0] - x[1]},
                    {"type": "eq", "fun": lambda x, *a: x[0] - x[1]})
        elif r < 0.45 and fn is not _pyargs:
            # a REGISTERED constraint set: fused with Sphere, generic route with every other objective
            cons = ({"type": "ineq", "fun": F.sum_le_6},)
        ref = m.shgo(fn, bounds, constraints=cons, **kw)
        res = shgo_b200.shgo(fn, bounds, constraints=cons, **kw)
        what = (seed, trial, getattr(fn, "__name__", "?"), bounds, kw, cons is not None)
        assert (res.success, res.nfev, res.nit, res.message) == (ref.success, ref.nfev, ref.nit, ref.message), what
        a, b = np.sort(np.atleast_1d(getattr(res, "funl", []))), np.sort(np.atleast_1d(getattr(ref, "funl", [])))
        assert len(a) == len(b), what
        np.testing.assert_allclose(a, b, rtol=1e-6, atol=1e-8, err_msg=str(what))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_problems_match_iteration_by_iteration(monkeypatch, seed):
    """The same differential idea at the level of the complex: after every `iterate()` the device
    complex and the driver's Python complex agree on vertex count, nfev, the pool (as a set; in its
    exact order wherever the complex is a mesh or a partial generation), the lowest vertex, the
    number of local minima found so far and the neighbour sets of every pool vertex."""
    _iterations_campaign(monkeypatch, seed, 15)


def _iterations_campaign(monkeypatch, seed, trials):
    fake_hotpath.install(monkeypatch)
    import shgo_b200
    from shgo_b200 import api, functors as F
    m = api.driver_module()
    funs = [F.rastrigin, F.ackley, F.styblinski_tang, F.sphere, _pyfun]
    rng = np.random.default_rng(2000 + seed)
    for trial in range(trials):
        d = int(rng.integers(1, 4))
        fn = funs[int(rng.integers(0, len(funs)))]
        lo = np.round(rng.uniform(-6, 0, d), 1)
        hi = lo + np.round(rng.uniform(0.5, 8, d), 1) + 0.5
        bounds = list(zip(lo.tolist(), hi.tolist()))
        method = ["simplicial", "sobol", "halton"][int(rng.integers(0, 3))]
        kw = dict(sampling_method=method, iters=3)
        if method == "simplicial":
            kw["n"] = [None, int(rng.integers(1, 40)), 2 ** d + 1][int(rng.integers(0, 3))]
        else:
            kw["n"] = int(rng.integers(4, 64))
        cons = None
        if rng.random() < 0.3 and d >= 2:
            c = float(rng.uniform(-2, 6))
            cons = ({"type": "ineq", "fun": lambda x, c=c: c - x[0] - x[1]},)
        what = (seed, trial, getattr(fn, "__name__", "?"), bounds, kw, cons is not None)
        r = m.SHGO(fn, bounds, constraints=cons, **kw)
        s = shgo_b200.SHGO(fn, bounds, constraints=cons, **kw)
        try:
            for it in range(3):
                r.iterate()
                s.iterate()
                assert (r.HC.V.size(), r.HC.V.nfev) == (s.HC.V.size(), s.HC.V.nfev), (what, it)
                rx = [tuple(np.round(x, 12)) for x in np.atleast_2d(r.X_min)] if len(r.X_min) else []
                sx = [tuple(np.round(x, 12)) for x in np.atleast_2d(s.X_min)] if len(s.X_min) else []
                assert sorted(rx) == sorted(sx), (what, it)
                if s.HC.mode != "lattice":
                    assert rx == sx, (what, it, s.HC.mode)
                r.find_lowest_vertex()
                s.find_lowest_vertex()
                assert (r.f_lowest is None) == (s.f_lowest is None), (what, it)
                if r.f_lowest is not None:
                    assert np.isclose(r.f_lowest, s.f_lowest, rtol=1e-9, atol=1e-12), (what, it)
                assert len(r.LMC.xl_maps) == len(s.LMC.xl_maps), (what, it)
                for v in getattr(s, "minimizer_pool", []):
                    rv = r.HC.V.cache[v.x]
                    assert {w.x for w in rv.nn if w is not rv} == {w.x for w in v.nn if w is not v}, (what, it, v.x)
        finally:
            api.uninstall()
