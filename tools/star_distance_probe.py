"""Star kernel time vs gather distance: neighbours i ^ 2^(j mod J) (degree fixed at log2 n), random f.
J small = every gather lands near the row; J = log2 n = the bench graph."""
import os, sys, torch
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
g = torch.Generator(device=dev); g.manual_seed(1)
f = torch.rand(n, dtype=torch.float64, device=dev, generator=g)
is_min = torch.empty(n, dtype=torch.uint8, device=dev)
row_ptr = torch.arange(0, (n + 1) * m, m, dtype=torch.int64, device=dev)
col = torch.empty((n, m), dtype=torch.int32, device=dev)
def build(bits_list):
    bits = torch.tensor([1 << b for b in bits_list], dtype=torch.int32, device=dev)[None, :]
    for s in range(0, n, 1 << 24):
        e = min(n, s + (1 << 24))
        i = torch.arange(s, e, dtype=torch.int32, device=dev)[:, None]
        torch.bitwise_xor(i, bits, out=col[s:e])
for J in (8, 12, 14, 16, 17, 18, 19, 20, 22, 24, log2n):
    # first 12 slots always j = 0..11 (rounds 1 and 2 unchanged); slots 12.. cycle through 12..J-1
    if J <= 12:
        bl = [j % J for j in range(m)]
    else:
        bl = list(range(12)) + [12 + (j % (J - 12)) for j in range(m - 12)]
    build(bl)
    ms = timeit(lambda: hp.star_csr(row_ptr, col.reshape(-1), f, is_min=is_min))
    print("J=%2d  %.3f ms   pool %d" % (J, ms, int(is_min.sum())))
# only ONE far slot, at the end: how much does a single far gather at alive ~1/28 cost?
for far in (20, 26):
    bl = list(range(12)) + [12 + (j % 4) for j in range(m - 13)] + [far]
    build(bl)
    ms = timeit(lambda: hp.star_csr(row_ptr, col.reshape(-1), f, is_min=is_min))
    print("near + one far (bit %d) last: %.3f ms" % (far, ms))
