"""Phase 2 of the sharded star test with the 'remote' shards living on THIS GPU: what the candidate
pass costs on the DRAM side alone (no NVLink).  Shard = 2^25 rows of the 2^28 bench graph, shard 3 of 8."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from shgo_b200 import hotpath as hp, _lib as L
log2n, G, s = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (28, 8, 3)))
n, m = 1 << log2n, log2n
rows = n // G
row0 = s * rows
dev = torch.device("cuda")
g = torch.Generator(device=dev); g.manual_seed(1)
f = torch.rand(n, dtype=torch.float64, device=dev, generator=g)
table = torch.tensor([f.data_ptr() + 8 * rows * k for k in range(G)], dtype=torch.int64, device=dev)
f_own = f[row0:row0 + rows]
rp = torch.arange(row0 * m, (row0 + rows + 1) * m, m, dtype=torch.int64, device=dev)
col = torch.empty((rows, m), dtype=torch.int32, device=dev)
bits = (1 << torch.arange(m, dtype=torch.int32, device=dev))[None, :]
for a in range(0, rows, 1 << 22):
    b = min(rows, a + (1 << 22))
    torch.bitwise_xor(torch.arange(row0 + a, row0 + b, dtype=torch.int32, device=dev)[:, None], bits, out=col[a:b])
col = col.reshape(-1)
flags = torch.empty(rows, dtype=torch.uint8, device=dev)
nfev = torch.zeros(1, dtype=torch.int64, device=dev)
cws = torch.empty(L.lib().shgo_b200_star_local_workspace(rows), dtype=torch.uint8, device=dev)
off0 = row0 * m
ENDS = [True]
def p1(): hp.star_csr_local(rp, col, row0, f_own, None, flags, nfev, off0, row0, rows, cws, ENDS[0])
def p2(ends): hp.star_csr_remote(rp, col, row0, table, f_own.data_ptr(), G, rows, s, flags, None, off0, row0, rows, ends, cws)
def t(fn, k=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / k
ENDS[0] = False
checked = t(p1)
ref_flags = flags.clone()
ENDS[0] = True
a = t(p1)
assert bool(((flags == ref_flags) | ((flags == 2) & (ref_flags == 0))).all())
print("phase 1: per-neighbour ownership test %.3f ms, rows trimmed at the ends %.3f ms" % (checked, a))
if G == 1:   # same rows through the single-GPU kernel: what the LOCAL bookkeeping costs by itself
    out = torch.empty(rows, dtype=torch.uint8, device=dev)
    print("single-GPU kernel on the same rows: %.3f ms" % t(lambda: hp.star_csr(rp, col, f, is_min=out)))
ncand = int((flags == 2).sum())
both = t(lambda: (p1(), p2(True)))
walk = t(lambda: (p1(), p2(False)))
print("rows %d  candidates %d (1/%.1f)  phase 1 %.3f ms  phase 2 (ends) %.3f ms  phase 2 (walk) %.3f ms"
      % (rows, ncand, rows / max(ncand, 1), a, both - a, walk - a))
