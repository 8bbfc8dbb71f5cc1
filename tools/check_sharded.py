"""Multi-GPU parity check (run under torchrun, one rank per GPU):
pool / nfev of the sharded iteration with the fused exchange == the two-phase p2p exchange == the
all_gather exchange == the single-GPU result computed by rank 0 on the whole problem.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/check_sharded.py [log2n]
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from shgo_b200 import dist as sd  # noqa: E402
from shgo_b200 import functors as F  # noqa: E402
from shgo_b200 import hotpath as hp  # noqa: E402

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n, m, d = 1 << log2n, log2n, 6
tab = hp.SobolTables([(-5.12, 5.12)] * d, dev)
row0, rows = sd.shard_range(n, world, rank)
rp_l = torch.arange(row0 * m, (row0 + rows + 1) * m, m, dtype=torch.int64, device=dev)
i = torch.arange(row0, row0 + rows, dtype=torch.int32, device=dev)[:, None]
col_l = (i ^ (1 << torch.arange(m, dtype=torch.int32, device=dev))[None, :]).reshape(-1)
res = {}
for ex in ("fused", "p2p", "allgather", "p2p-walk", "p2p-nofilter"):
    it = sd.ShardedSobolIteration(F.F_RASTRIGIN, F.G_NONE, tab, n, rp_l, col_l, exchange=ex.split("-")[0],
                                  remote_at_ends=(ex in ("p2p", "fused", "p2p-nofilter")),
                                  filter_remote=(ex != "p2p-nofilter"))
    assert it.filter_remote == (ex in ("p2p", "p2p-walk")) if hasattr(it, "filter_remote") else True
    assert it.exchange == ex.split("-")[0]
    for _ in range(3):
        it.step()
    torch.cuda.synchronize()
    res[ex] = (it.global_pool(), it.global_nfev())
    it.close()
ok = np.array_equal(res["p2p"][0], res["allgather"][0]) and res["p2p"][1] == res["allgather"][1]
ok = ok and np.array_equal(res["p2p"][0], res["p2p-walk"][0])
ok = ok and np.array_equal(res["p2p"][0], res["fused"][0]) and res["p2p"][1] == res["fused"][1]
ok = ok and np.array_equal(res["p2p"][0], res["p2p-nofilter"][0]) and res["p2p"][1] == res["p2p-nofilter"][1]
if rank == 0:
    rp, col = hp.hypercube_csr(n, dev)
    f, feas = hp.eval_sobol(F.F_RASTRIGIN, F.G_NONE, tab, 0, n)
    idx, cnt = hp.compact_pool(hp.star_csr(rp, col, f))
    single = idx[: int(cnt.item())].cpu().numpy()
    ok = ok and np.array_equal(single, res["p2p"][0]) and res["p2p"][1] == n
    print("check_sharded world=%d n=2^%d pool=%d nfev=%d : %s" % (world, log2n, len(single), res["p2p"][1],
                                                                    "OK" if ok else "MISMATCH"))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
