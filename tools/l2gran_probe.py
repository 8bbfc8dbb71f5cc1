"""Does the L2 fetch granularity hint (cudaLimitMaxL2FetchGranularity) change the star kernel's
gather over-fetch?  Times the bench graph and a locality-free graph at 32 / 64 / 128 bytes."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from shgo_b200 import hotpath as hp
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 27
n, m = 1 << log2n, log2n
dev = torch.device("cuda")
torch.zeros(1, device=dev)
rt = None
for name in ("libcudart.so.12", "libcudart.so"):
    try:
        rt = ctypes.CDLL(name); break
    except OSError:
        pass
LIMIT = 0x05  # cudaLimitMaxL2FetchGranularity
def get():
    v = ctypes.c_size_t(0)
    rc = rt.cudaDeviceGetLimit(ctypes.byref(v), LIMIT)
    return rc, v.value
def timeit(fn, k=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / k
row_ptr, col = hp.hypercube_csr(n, dev)
g = torch.Generator(device=dev); g.manual_seed(1)
f = torch.rand(n, dtype=torch.float64, device=dev, generator=g)
colr = torch.randint(0, n, (n * m,), dtype=torch.int32, device=dev, generator=g)
is_min = torch.empty(n, dtype=torch.uint8, device=dev)
print("default limit:", get())
for gran in (None, 32, 64, 128, 32):
    if gran is not None:
        rc = rt.cudaDeviceSetLimit(LIMIT, ctypes.c_size_t(gran))
        print("set", gran, "rc", rc, "now", get())
    a = timeit(lambda: hp.star_csr(row_ptr, col, f, is_min=is_min))
    b = timeit(lambda: hp.star_csr(row_ptr, colr, f, is_min=is_min), k=2)
    print("gran %s: hypercube %.3f ms   random-graph %.3f ms" % (gran, a, b))
