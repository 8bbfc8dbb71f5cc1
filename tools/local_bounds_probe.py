"""Local bounds of pool vertices: device kernels vs the reference's Python walk."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import shgo_b200
from shgo_b200 import api, functors as F, hotpath as hp, lattice

def dev_time(fn, k=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / k

# 1. reference-sized case: 7-D, generation 1, through the driver seam on both sides
d, bounds = 7, [(-5.0, 4.0)] * 7
s = shgo_b200.SHGO(F.styblinski_tang, bounds, n=None, iters=2, sampling_method="simplicial")
s.iterate_complex(); s.iterate_complex(); s.minimizers()
pool = list(s.minimizer_pool)
t0 = time.perf_counter(); boxes = [s.construct_lcb_simplicial(v) for v in pool]; t_dev = time.perf_counter() - t0
api.uninstall()
m = api.driver_module()
r = m.SHGO(F.styblinski_tang, bounds, n=None, iters=2, sampling_method="simplicial")
r.iterate_complex(); r.iterate_complex(); r.minimizers()
rpool = list(r.minimizer_pool)
t0 = time.perf_counter(); rboxes = [r.construct_lcb_simplicial(v) for v in rpool]; t_ref = time.perf_counter() - t0
key = lambda bs: sorted(tuple(np.asarray(b, float).ravel().tolist()) for b in bs)
same = key(boxes) == key(rboxes)
print("7-D gen 1: |V|=%d pool=%d  device seam %.2f ms (all boxes, incl. launch + readback)  "
      "python Complex %.2f ms  identical=%s" % (s.HC.V.size(), len(pool), t_dev * 1e3, t_ref * 1e3, same))

# 2. kernel alone at scale: 10-D generation 2 (10.8 M vertices), pool = every 50th vertex
d, gen = 10, 2
b10 = [(-32.768, 32.768)] * d
tab = torch.from_numpy(lattice.coordinate_tables(b10, gen)).cuda()
nv, _ = lattice.counts(d, gen)
pool = torch.arange(0, nv, 50, dtype=torch.int64).cuda()
ms = dev_time(lambda: hp.lattice_local_bounds(d, gen, tab, pool, b10))
print("lattice 10-D gen 2: pool %d -> %.3f ms (%.1f M boxes/s)" % (pool.numel(), ms, pool.numel() / ms / 1e3))

# 3. explicit CSR: 2^24-vertex hypercube graph in 8-D Sobol coordinates, pool = star minimisers
n = 1 << 24
st = hp.SobolTables([(-5.0, 5.0)] * 8)
x = hp.sobol_points(st, 0, n, "soa")
rp, col = hp.hypercube_csr(n)
f, _ = hp.eval_sobol(F.functor_id(F.styblinski_tang), 0, st, 0, n)
is_min = hp.star_csr(rp, col, f)
idx, cnt = hp.compact_pool(is_min)
pool = idx[: int(cnt.item())]
ms = dev_time(lambda: hp.local_bounds_csr(rp, col, x, pool, [(-5.0, 5.0)] * 8))
print("CSR 2^24 x 8-D, degree 24: pool %d -> %.3f ms (%.1f M boxes/s)" % (pool.numel(), ms, pool.numel() / ms / 1e3))
