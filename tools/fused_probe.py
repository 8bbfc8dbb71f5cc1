"""One GPU, shard 0 of `nshards` emulated (all shards in local memory): where does the fused star
kernel's time go?  local phase 1 / fused with every peer ready / fused with no peer ready / phase 2.

    python tools/fused_probe.py [log2 rows per shard] [nshards]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from shgo_b200 import _lib, hotpath as hp  # noqa: E402

lr = int(sys.argv[1]) if len(sys.argv) > 1 else 26
nsh = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rows = 1 << lr
n = rows * nsh
m = n.bit_length() - 1
dev = torch.device("cuda")
row_ptr = torch.arange(0, (rows + 1) * m, m, dtype=torch.int64, device=dev)
i = torch.arange(0, rows, dtype=torch.int32, device=dev)[:, None]
col = (i ^ (1 << torch.arange(m, dtype=torch.int32, device=dev))[None, :]).reshape(-1)
del i
g = torch.Generator(device=dev)
g.manual_seed(1)
f = torch.rand(n, dtype=torch.float64, device=dev, generator=g)
shards = [f[s * rows:(s + 1) * rows] for s in range(nsh)]
table = torch.tensor([t.data_ptr() for t in shards], dtype=torch.int64, device=dev)
flags = torch.empty(rows, dtype=torch.uint8, device=dev)
nfev = torch.zeros(1, dtype=torch.int64, device=dev)
cws = torch.empty(_lib.lib().shgo_b200_star_local_workspace(rows), dtype=torch.uint8, device=dev)
epoch = 5
ready = torch.full((64,), epoch, dtype=torch.int32, device=dev)
notready = torch.zeros(64, dtype=torch.int32, device=dev)


def timeit(fn, k=6):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / k


def local():
    hp.star_csr_local(row_ptr, col, 0, shards[0], None, flags, nfev, 0, cand_ws=cws, remote_at_ends=True)


def fused(rdy):
    hp.star_csr_fused(row_ptr, col, 0, table, shards[0], None, nsh, rows, 0, rdy, epoch, flags, cws, nfev, 0)


def remote():
    hp.star_csr_remote(row_ptr, col, 0, table, shards[0].data_ptr(), nsh, rows, 0, flags, None, 0,
                       remote_at_ends=True, cand_ws=cws)


t_local = timeit(local)
ncand = int((flags == 2).sum())
t_remote = timeit(lambda: (local(), remote())) - t_local
ref = flags.clone()
t_fr = timeit(lambda: fused(ready))
left = int((flags == 2).sum())
remote()
ok1 = torch.equal(flags, ref)
t_fn = timeit(lambda: fused(notready))
remote()
ok2 = torch.equal(flags, ref)
print("rows 2^%d x %d shards (deg %d): local %.3f ms (%d candidates) | phase 2 %.3f ms | fused, peers ready %.3f ms "
      "(%d left for phase 2) | fused, nobody ready %.3f ms | parity %s %s"
      % (lr, nsh, m, t_local, ncand, t_remote, t_fr, left, t_fn, ok1, ok2))
