"""Where does the star kernel's time go?  Same CSR volume, different gather locality / survival."""
import sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from shgo_b200 import hotpath as hp
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 27
n, m = 1 << log2n, log2n
dev = torch.device("cuda")
def timeit(fn, k=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / k
row_ptr, col = hp.hypercube_csr(n, dev)
g = torch.Generator(device=dev); g.manual_seed(1)
f_rand = torch.rand(n, dtype=torch.float64, device=dev, generator=g)
f_const = torch.zeros(n, dtype=torch.float64, device=dev)
f_ramp = torch.arange(n, dtype=torch.float64, device=dev)          # only row 0 survives everything
is_min = torch.empty(n, dtype=torch.uint8, device=dev)
alg = 4 * col.numel() + 17 * n
def report(name, rp, c, f):
    ms = timeit(lambda: hp.star_csr(rp, c, f, is_min=is_min))
    print("%-44s %7.3f ms  %6.0f GB/s algorithmic  pool %d" % (name, ms, alg / ms / 1e6, int(is_min.sum())))
report("hypercube, random f", row_ptr, col, f_rand)
report("hypercube, constant f (all fail in round 1)", row_ptr, col, f_const)
report("hypercube, ramp f", row_ptr, col, f_ramp)
# banded graph, same degree: neighbours r-m/2..r+m/2 (clamped by wrap-around), sorted ascending
offs = torch.cat([torch.arange(-(m // 2), 0), torch.arange(1, m - m // 2 + 1)]).to(dev).to(torch.int32)
colb = torch.empty((n, m), dtype=torch.int32, device=dev)
for s in range(0, n, 1 << 22):
    e = min(n, s + (1 << 22))
    i = torch.arange(s, e, dtype=torch.int32, device=dev)[:, None]
    colb[s:e] = (i + offs[None, :]) & (n - 1)
colb = colb.reshape(-1)
report("banded +-%d, random f" % (m // 2), row_ptr, colb, f_rand)
report("banded, constant f", row_ptr, colb, f_const)
# random graph (no locality at all), same degree
colr = torch.randint(0, n, (n * m,), dtype=torch.int32, device=dev, generator=g)
report("random neighbours, random f", row_ptr, colr, f_rand)
report("random neighbours, constant f", row_ptr, colr, f_const)
# pure streaming reference: read col + row_ptr with torch
ms = timeit(lambda: (col.sum(), row_ptr.sum()))
print("torch sum(col)+sum(row_ptr) %.3f ms  %.0f GB/s" % (ms, (4 * col.numel() + 8 * n) / ms / 1e6))
