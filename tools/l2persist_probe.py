"""Does a persisting-L2 set-aside (cudaLimitPersistingL2CacheSize) change how long the evict_last
f lines of the star kernel survive?  Times the bench graph with 0 / 16 / 32 / 64 MiB / max set aside."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from shgo_b200 import hotpath as hp
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 27
n = 1 << log2n
dev = torch.device("cuda")
torch.zeros(1, device=dev)
rt = ctypes.CDLL("libcudart.so.12")
LIMIT = 0x06  # cudaLimitPersistingL2CacheSize
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
props = torch.cuda.get_device_properties(0)
print("L2 size", props.L2_cache_size, "default persisting limit", get())
row_ptr, col = hp.hypercube_csr(n, dev)
g = torch.Generator(device=dev); g.manual_seed(1)
f = torch.rand(n, dtype=torch.float64, device=dev, generator=g)
is_min = torch.empty(n, dtype=torch.uint8, device=dev)
for mib in (None, 16, 32, 64, 96, 128, 0):
    if mib is not None:
        rc = rt.cudaDeviceSetLimit(LIMIT, ctypes.c_size_t(mib << 20))
        print("set %d MiB rc %d now %s" % (mib, rc, get()))
    print("  star %.3f ms" % timeit(lambda: hp.star_csr(row_ptr, col, f, is_min=is_min)))
