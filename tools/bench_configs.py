"""Secondary measurements for the other BASELINE.json configs (not the bench.py headline):
cfg2 Rastrigin 6-D 2^20 / cfg4 cattle-feed 4-D 2^22 (synthetic CSR), cfg3 Ackley 10-D lattice
generations 0-2, K3 CSR build from real qhull simplices, and whole `shgo()` calls against the
driver with its own Python Complex on the same host.  Prints a markdown table."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import shgo_b200  # noqa: E402
from shgo_b200 import api, functors as F, hotpath as hp, lattice  # noqa: E402


def timeit(fn, k=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / k


rows = []


def sobol_cfg(name, fid, gid, bounds, log2n):
    n = 1 << log2n
    tab = hp.SobolTables(bounds)
    rp, col = hp.hypercube_csr(n)
    f = torch.empty(n, dtype=torch.float64, device="cuda")
    feas = torch.empty(n, dtype=torch.uint8, device="cuda")
    is_min = torch.empty(n, dtype=torch.uint8, device="cuda")
    nf = torch.zeros(1, dtype=torch.int64, device="cuda")
    work = hp.PoolWorkspace(n, n, torch.device("cuda"))
    t_eval = timeit(lambda: hp.eval_sobol(fid, gid, tab, 0, n, f=f, feasible=feas))
    t_star = timeit(lambda: hp.star_csr(rp, col, f, is_min=is_min, feasible=feas, nfev=nf))
    t_pool = timeit(lambda: hp.compact_pool(is_min, work=work))
    tot = t_eval + t_star + t_pool
    rows.append("| %s | %d | %.3f | %.3f | %.3f | %.3g | feasible %d, pool %d |" % (
        name, n, t_eval, t_star, t_pool, n / tot * 1e3, int(nf.item()), int(work.count[0].item())))


def lattice_cfg(d, gen, fid, bounds):
    tabh = lattice.coordinate_tables(bounds, gen)
    tab = torch.from_numpy(tabh).cuda()
    nv, _ = lattice.counts(d, gen)
    f = torch.empty(nv, dtype=torch.float64, device="cuda")
    feas = torch.empty(nv, dtype=torch.uint8, device="cuda")
    is_min = torch.empty(nv, dtype=torch.uint8, device="cuda")
    k = 10 if nv < 1e6 else 3
    t_eval = timeit(lambda: hp.lattice_eval(fid, 0, d, gen, tab, 0, nv, f=f, feasible=feas), k)
    t_star = timeit(lambda: hp.lattice_star(d, gen, f, is_min=is_min), k)
    rows.append("| Ackley %d-D lattice generation %d | %d | %.3f | %.3f | - | %.3g | minimisers %d |" % (
        d, gen, nv, t_eval, t_star, nv / (t_eval + t_star) * 1e3, int(is_min.sum().item())))


def csr_build(d, log2n):
    from scipy.spatial import Delaunay
    from scipy.stats import qmc
    n = 1 << log2n
    pts = qmc.Sobol(d=d, scramble=False, seed=0).random(n)
    t0 = time.perf_counter()
    tri = Delaunay(pts, incremental=True)
    t_qhull = time.perf_counter() - t0
    simp = torch.from_numpy(tri.simplices.astype(np.int32)).cuda()
    t = timeit(lambda: hp.csr_from_simplices(simp, n), 5)
    rp, col = hp.csr_from_simplices(simp, n)
    rows.append("| K3 vf_to_vv, %d-D Sobol n=2^%d, %d simplices (qhull %.1f s on host) | %d | - | - | - | %.3g simplices/s | %.3f ms, nnz %d |" % (
        d, log2n, len(tri.simplices), t_qhull, n, len(tri.simplices) / t * 1e3, t, col.numel()))


def hot_path_in_driver(name, fn, bounds, n):
    """The hot path as the driver sees it (everything of iterate_delaunay after qhull + minimizers):
    vf_to_vv + process_pools + minimizers on the SAME simplices, Python Complex vs DeviceComplex."""
    m = api.driver_module()
    times = {}
    for label in ("python", "device"):
        if label == "device":
            s_ = shgo_b200.SHGO(fn, bounds, n=n, iters=1, sampling_method="sobol")
        else:
            api.uninstall()
            s_ = m.SHGO(fn, bounds, n=n, iters=1, sampling_method="sobol")
        s_.nc += s_.n
        s_.sampled_surface(infty_cons_sampl=s_.infty_cons_sampl)
        s_.delaunay_triangulation(n_prc=0)
        best = None
        for rep in range(2 if label == "device" else 1):
            if rep:       # fresh complex for the repeat (first device pass pays cudaMalloc / module load)
                s_.HC = type(s_.HC)(dim=s_.dim, domain=s_.bounds, sfield=s_.func, constraints=s_.constraints)
                s_.HC._sampler = getattr(s_, "_b200_sampler", None)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            s_.HC.vf_to_vv(s_.Tri.points, s_.Tri.simplices)
            s_.HC.V.process_pools()
            s_.minimizers()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        times[label] = (best, len(s_.minimizer_pool), s_.HC.V.nfev, len(s_.Tri.simplices))
    api.uninstall()
    (tp, pp, nfp, S), (td, pd, nfd, _) = times["python"], times["device"]
    rows.append("| driver hot path (vf_to_vv + process_pools + minimizers) %s | %d pts, %d simplices | Python Complex %.3f s | DeviceComplex %.4f s | - | x%.0f | pool %d/%d, nfev %d/%d |" % (
        name, n, S, tp, td, tp / td, pp, pd, nfp, nfd))


def whole_call(name, fn, bounds, **kw):
    m = api.driver_module()
    t0 = time.perf_counter()
    ref = m.shgo(fn, bounds, **kw)
    t_ref = time.perf_counter() - t0
    shgo_b200.shgo(fn, bounds, **kw)        # warm-up (module load, cudaMalloc)
    t0 = time.perf_counter()
    res = shgo_b200.shgo(fn, bounds, **kw)
    t_dev = time.perf_counter() - t0
    same = (res.nfev == ref.nfev and len(res.funl) == len(ref.funl)
            and np.allclose(res.funl, ref.funl, rtol=1e-7, atol=1e-9))
    rows.append("| whole `shgo()` %s | nfev %d | driver+Python Complex %.2f s | driver+DeviceComplex %.2f s | - | x%.1f | same result: %s, %d minima |" % (
        name, res.nfev, t_ref, t_dev, t_ref / t_dev, same, len(res.funl)))


ONLY = os.environ.get("ONLY", "")
if ONLY == "driver":
    hot_path_in_driver("Eggholder 2-D", F.eggholder, [(-512, 512)] * 2, 1 << 16)
    hot_path_in_driver("Rastrigin 3-D", F.rastrigin, [(-5.12, 5.12)] * 3, 1 << 14)
    hot_path_in_driver("Rastrigin 5-D", F.rastrigin, [(-5.12, 5.12)] * 5, 1 << 10)
    print("\n".join(rows))
    sys.exit(0)
if ONLY == "lattice":
    for g in (0, 1, 2):
        lattice_cfg(10, g, F.F_ACKLEY, [(-32.768, 32.768)] * 10)
    lattice_cfg(8, 2, F.F_ACKLEY, [(-32.768, 32.768)] * 8)
    lattice_cfg(6, 4, F.F_ACKLEY, [(-32.768, 32.768)] * 6)
    print("\n".join(rows))
    sys.exit(0)
sobol_cfg("cfg2 Rastrigin 6-D, Sobol 2^20, hypercube CSR", F.F_RASTRIGIN, 0, [(-5.12, 5.12)] * 6, 20)
sobol_cfg("cfg4 cattle-feed 4-D + g1,g2, Sobol 2^22, hypercube CSR", F.F_CATTLE_FEED, F.G_CATTLE_FEED, [(0, 1.0)] * 4, 22)
sobol_cfg("cfg5 Styblinski-Tang 8-D, Sobol 2^24, hypercube CSR", F.F_STYBLINSKI_TANG, 0, [(-5, 5)] * 8, 24)
sobol_cfg("Ackley 10-D, Sobol 2^24, hypercube CSR", F.F_ACKLEY, 0, [(-32.768, 32.768)] * 10, 24)
for g in (0, 1, 2):
    lattice_cfg(10, g, F.F_ACKLEY, [(-32.768, 32.768)] * 10)
csr_build(2, 18)
csr_build(3, 15)
csr_build(4, 12)
hot_path_in_driver("Eggholder 2-D", F.eggholder, [(-512, 512)] * 2, 1 << 16)
hot_path_in_driver("Rastrigin 3-D", F.rastrigin, [(-5.12, 5.12)] * 3, 1 << 14)
hot_path_in_driver("Rastrigin 5-D", F.rastrigin, [(-5.12, 5.12)] * 5, 1 << 10)
whole_call("Eggholder 2-D, sobol n=2^14", F.eggholder, [(-512, 512)] * 2, n=1 << 14, sampling_method="sobol")
whole_call("Rastrigin 3-D, sobol n=2^12", F.rastrigin, [(-5.12, 5.12)] * 3, n=1 << 12, sampling_method="sobol")
whole_call("Ackley 4-D, simplicial n=None iters=3", F.ackley, [(-32.768, 32.768)] * 4, n=None, iters=3)
print("| case | n | eval ms | star ms | pool ms | points/s | notes |\n|---|---|---|---|---|---|---|")
print("\n".join(rows))
