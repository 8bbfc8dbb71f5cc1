"""Generate the golden fixtures in this directory from the LIVE, UNMODIFIED reference.

Run in the build container only (needs /root/reference; the GPU box has no reference):

    python tests/golden/make_golden.py

Recipe = SURVEY.md Appendix C: drive the reference one stage at a time and record what the
hot path produced: the points handed to qhull, the simplices qhull returned, the vertex graph
`vf_to_vv` built, f per vertex, feasibility, nfev and the minimiser-pool index set.  The
objectives are the numpy callables of oracle/oracle.py (the reference only defines Eggholder,
cattle-feed and, through scipy, Rosenbrock itself).

Everything written here is deterministic (unscrambled Sobol, no seeds); qhull's simplex list
depends on the scipy build (1.18.1 here), which is why the simplices are part of the fixture.
"""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))

from shgo import SHGO, shgo  # noqa: E402  (the reference)
from shgo._shgo_lib._complex import Complex  # noqa: E402
from scipy.optimize import rosen  # noqa: E402

from oracle import oracle as orc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def small_int(a):
    a = np.asarray(a)
    for dt in (np.uint8, np.uint16, np.int32):
        if a.size == 0 or a.max() <= np.iinfo(dt).max:
            return a.astype(dt)
    return a


def sobol_case(name, functor, gset, bounds, n, full_simplices=True):
    fn = orc.NP_OBJECTIVES[functor]
    cons = tuple({"type": "ineq", "fun": g} for g in orc.NP_CONSTRAINTS[gset]) or None
    s = SHGO(fn, bounds, n=n, iters=1, constraints=cons, sampling_method="sobol")
    s.iterate_complex()
    pts, simp = s.Tri.points, s.Tri.simplices
    idx = {tuple(p): i for i, p in enumerate(pts)}
    assert len(idx) == len(pts)
    edges = sorted({(min(idx[x], idx[w.x]), max(idx[x], idx[w.x]))
                    for x, v in s.HC.V.cache.items() for w in v.nn})
    present = np.zeros(len(pts), np.uint8)
    f = np.full(len(pts), np.nan)
    feas = np.zeros(len(pts), np.uint8)
    for x, v in s.HC.V.cache.items():
        present[idx[x]] = 1
        f[idx[x]] = v.f
        feas[idx[x]] = 1 if (cons is None or v.feasible) else 0
    s.minimizers()
    pool = np.array(sorted(idx[tuple(v.x)] for v in s.minimizer_pool), np.int64)
    h = hashlib.sha256(pool.tobytes()).hexdigest()[:16]
    dp1 = simp.shape[1]
    simp_store = simp if full_simplices else simp[:, :3]
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        functor=functor, gset=str(gset), bounds=np.array(bounds, float), n=n,
        points=pts, simplices=small_int(simp_store), dp1=dp1,
        edges=small_int(np.array(edges).reshape(-1, 2)), present=present, f=f, feasible=feas,
        pool=pool, nfev=s.HC.V.nfev, pool_hash=h)
    print(f"{name}: S={len(simp)} |V|={int(present.sum())} nfev={s.HC.V.nfev} "
          f"pool={len(pool)} first={pool[:6].tolist()} hash={h}")


def lattice_case(name, bounds, gens):
    """Vertex coordinates (cache order) and edge list of the reference's hypercube complex
    after triangulate() + g x refine_all()."""
    d = len(bounds)
    HC = Complex(d, domain=bounds, sfield=lambda x: 0.0)
    out = {}
    for g in range(gens + 1):
        if g == 0:
            HC.triangulate()
        else:
            HC.refine_all()
        keys = list(HC.V.cache.keys())
        idx = {k: i for i, k in enumerate(keys)}
        edges = sorted({(min(idx[k], idx[w.x]), max(idx[k], idx[w.x]))
                        for k, v in HC.V.cache.items() for w in v.nn})
        out[f"coords_g{g}"] = np.array(keys, float).reshape(len(keys), d)
        out[f"edges_g{g}"] = small_int(np.array(edges).reshape(-1, 2))
        print(f"{name} g{g}: |V|={len(keys)} |E|={len(edges)}")
    np.savez_compressed(os.path.join(HERE, name + ".npz"), bounds=np.array(bounds, float),
                        gens=gens, **out)


def simplicial_pools_case(name, functor, gset, bounds, iters):
    """Whole-generation simplicial run (n=None): per-iteration minimiser pools (coordinates,
    recorded before minimise_pool trims them), vertex counts and nfev."""
    fn = orc.NP_OBJECTIVES[functor]
    cons = tuple({"type": "ineq", "fun": g} for g in orc.NP_CONSTRAINTS[gset]) or None
    s = SHGO(fn, bounds, n=None, iters=iters, constraints=cons, sampling_method="simplicial")
    s.n = None  # whole generations (refine_all) -- _shgo.py:1015-1017
    rec = {}
    orig = s.minimise_pool
    it = [0]

    def spy(force_iter=False):
        rec[f"pool_x_{it[0]}"] = np.array(s.X_min, float)
        rec[f"pool_f_{it[0]}"] = np.array(s.minimizer_pool_F, float)
        return orig(force_iter)

    s.minimise_pool = spy
    while not s.stop_global:
        s.iterate()
        rec[f"nvert_{it[0]}"] = s.HC.V.size()
        rec[f"nfev_{it[0]}"] = s.HC.V.nfev
        if f"pool_x_{it[0]}" not in rec:
            rec[f"pool_x_{it[0]}"] = np.zeros((0, len(bounds)))
            rec[f"pool_f_{it[0]}"] = np.zeros(0)
        print(f"{name} it{it[0]}: |V|={s.HC.V.size()} nfev={s.HC.V.nfev} "
              f"pool={len(rec[f'pool_x_{it[0]}'])}")
        it[0] += 1
        s.stopping_criteria()
    s.find_minima() if not s.minimize_every_iter else None
    np.savez_compressed(os.path.join(HERE, name + ".npz"), functor=functor, gset=str(gset),
                        bounds=np.array(bounds, float), iters=iters, **rec)


def _lcb_of(HC, bounds):
    """construct_lcb_simplicial (_shgo.py:1246-1278) of EVERY vertex of HC, through the
    reference's own method (unbound call on a stand-in carrying the two attributes it reads)."""
    class _Self:
        pass
    me = _Self()
    me.bounds, me.disp = bounds, False
    keys = list(HC.V.cache.keys())
    idx = {k: i for i, k in enumerate(keys)}
    lo = np.empty((len(keys), len(bounds)))
    hi = np.empty((len(keys), len(bounds)))
    for k, v in HC.V.cache.items():
        cb = SHGO.construct_lcb_simplicial(me, v)
        lo[idx[k]] = [c[0] for c in cb]
        hi[idx[k]] = [c[1] for c in cb]
    edges = sorted({(min(idx[k], idx[w.x]), max(idx[k], idx[w.x]))
                    for k, v in HC.V.cache.items() for w in v.nn})
    return (np.array(keys, float).reshape(len(keys), len(bounds)),
            small_int(np.array(edges).reshape(-1, 2)), lo, hi)


def local_bounds_cases():
    """local bounds of every vertex, for whole generations of the
    hypercube complex, a partial refine(n) state and a Delaunay mesh."""
    out = {}
    names = []

    def put(name, bounds, HC):
        x, e, lo, hi = _lcb_of(HC, bounds)
        out[name + "_bounds"] = np.array(bounds, float)
        out[name + "_x"], out[name + "_edges"] = x, e
        out[name + "_lo"], out[name + "_hi"] = lo, hi
        names.append(name)
        print(f"local_bounds {name}: |V|={len(x)} |E|={len(e)}")

    for name, bounds, gens in (("lat2d", [(0.0, 1.0)] * 2, 3),
                               ("lat3d_mixed", [(-3.3, 7.1), (0.0, 2.0), (-5.0, 5.0)], 2),
                               ("lat4d", [(-5.12, 5.12)] * 4, 1)):
        HC = Complex(len(bounds), domain=bounds, sfield=lambda x: 0.0)
        HC.triangulate()
        for g in range(gens + 1):
            if g:
                HC.refine_all()
            put(f"{name}_g{g}", bounds, HC)
    bounds = [(-1.0, 2.0)] * 3
    HC = Complex(3, domain=bounds, sfield=lambda x: 0.0)
    HC.refine(9)
    HC.refine(17)
    HC.refine(30)
    put("partial3d", bounds, HC)
    bounds = [(-5.12, 5.12)] * 3
    s = SHGO(orc.np_rastrigin, bounds, n=256, iters=1, sampling_method="sobol")
    s.iterate_complex()
    put("mesh3d", bounds, s.HC)
    np.savez_compressed(os.path.join(HERE, "local_bounds.npz"), names=np.array(names), **out)


def end_to_end():
    out = {}
    r = shgo(rosen, [(0, 2)] * 5)
    out["rosen5_x"], out["rosen5_fun"], out["rosen5_nfev"] = r.x, r.fun, r.nfev
    print("README example:", r.x, r.fun, r.nfev)
    cons = ({"type": "ineq", "fun": orc.np_cattle_g1}, {"type": "ineq", "fun": orc.np_cattle_g2},
            {"type": "eq", "fun": lambda x: x[0] + x[1] + x[2] + x[3] - 1})
    r = shgo(orc.np_cattle_feed, [(0, 1.0)] * 4, n=150, constraints=cons)
    out["cattle_x"], out["cattle_fun"], out["cattle_nfev"] = r.x, r.fun, r.nfev
    print("cattle-feed:", r.x, r.fun, r.nfev, r.nlfev)
    r = shgo(orc.np_eggholder, [(-512, 512)] * 2, n=64, sampling_method="sobol")
    out["egg_xl"], out["egg_funl"], out["egg_nfev"] = r.xl, r.funl, r.nfev
    print("eggholder sobol n=64:", len(r.funl), r.fun, r.nfev)
    np.savez_compressed(os.path.join(HERE, "end_to_end.npz"), **out)


if __name__ == "__main__":
    if sys.argv[1:] == ["local_bounds"]:      # fixtures are independent: regenerate just this one
        local_bounds_cases()
        sys.exit(0)
    if sys.argv[1:] == ["simplicial_10d"]:    # ten minutes: the reference builds 10-D generation 1
        simplicial_pools_case("simplicial_ackley_10d_it2", "ackley", None, [(-32.768, 32.768)] * 10, 2)
        sys.exit(0)
    if sys.argv[1:] == ["lattice_high_d"]:
        lattice_case("lattice_unit_6d", [(0.0, 1.0)] * 6, 1)
        lattice_case("lattice_unit_7d", [(0.0, 1.0)] * 7, 1)
        sys.exit(0)
    B512 = [(-5.12, 5.12)]
    sobol_case("sobol_rastrigin_6d_256", "rastrigin", None, B512 * 6, 256, full_simplices=False)
    sobol_case("sobol_rastrigin_3d_4096", "rastrigin", None, B512 * 3, 4096)
    sobol_case("sobol_cattle_4d_1024", "cattle_feed", "cattle_feed", [(0, 1.0)] * 4, 1024)
    sobol_case("sobol_cattle_4d_4096", "cattle_feed", "cattle_feed", [(0, 1.0)] * 4, 4096,
               full_simplices=False)
    sobol_case("sobol_styblinski_8d_128", "styblinski_tang", None, [(-5, 5)] * 8, 128,
               full_simplices=False)
    sobol_case("sobol_styblinski_3d_4096", "styblinski_tang", None, [(-5, 5)] * 3, 4096)
    sobol_case("sobol_eggholder_2d_4096", "eggholder", None, [(-512, 512)] * 2, 4096)
    sobol_case("sobol_ackley_3d_2048", "ackley", None, [(-32.768, 32.768)] * 3, 2048)
    sobol_case("sobol_rosenbrock_4d_512", "rosenbrock", None, [(0, 2)] * 4, 512)
    sobol_case("sobol_sphere_2d_256_cons", "sphere", "sum_le_6", [(-1, 6)] * 2, 256)
    sobol_case("sobol_sphere_1d_64", "sphere", None, [(-1, 6)], 64)

    lattice_case("lattice_unit_2d", [(0.0, 1.0)] * 2, 3)
    lattice_case("lattice_unit_3d", [(0.0, 1.0)] * 3, 2)
    lattice_case("lattice_rastrigin_4d", B512 * 4, 2)
    lattice_case("lattice_rosen_5d", [(0.0, 2.0)] * 5, 1)
    lattice_case("lattice_ackley_3d", [(-32.768, 32.768)] * 3, 2)
    lattice_case("lattice_mixed_3d", [(-3.3, 7.1), (0.0, 2.0), (-5.0, 5.0)], 2)
    lattice_case("lattice_unit_1d", [(0.0, 1.0)], 3)
    lattice_case("lattice_unit_6d", [(0.0, 1.0)] * 6, 1)       # 793 / 18 928
    lattice_case("lattice_unit_7d", [(0.0, 1.0)] * 7, 1)       # 2315 / 92 194

    simplicial_pools_case("simplicial_ackley_3d_it3", "ackley", None, [(-32.768, 32.768)] * 3, 3)
    simplicial_pools_case("simplicial_rastrigin_2d_it4", "rastrigin", None, B512 * 2, 4)
    simplicial_pools_case("simplicial_styblinski_4d_it3", "styblinski_tang", None,
                          [(-5, 5)] * 4, 3)
    simplicial_pools_case("simplicial_sphere_2d_cons_it3", "sphere", "sum_le_6",
                          [(-1, 6)] * 2, 3)
    simplicial_pools_case("simplicial_ackley_10d_it2", "ackley", None, [(-32.768, 32.768)] * 10, 2)
    end_to_end()
    local_bounds_cases()
