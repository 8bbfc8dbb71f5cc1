"""Multi-GPU simplicial lattice: parity and timing, run under torchrun.

pool / nfev of ShardedLatticeIteration with the push exchange == with the all_reduce exchange == the
single-GPU result computed by rank 0 on the whole generation; then ms per generation (max over
ranks, CUDA events) for both exchanges and for one GPU alone.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29513 tools/check_sharded_lattice.py [d gen]
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from shgo_b200 import dist as sd  # noqa: E402
from shgo_b200 import functors as F  # noqa: E402
from shgo_b200 import hotpath as hp  # noqa: E402
from shgo_b200 import lattice  # noqa: E402

d = int(sys.argv[1]) if len(sys.argv) > 1 else 10
gen = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
bounds = [(-32.768, 32.768)] * d
tab = torch.from_numpy(lattice.coordinate_tables(bounds, gen)).to(dev)
nv, ngrid = lattice.counts(d, gen)
fid, gid = F.F_ACKLEY, F.G_NONE


def timed(step, k=10, w=3):
    for _ in range(w):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(k):
        step()
    e1.record()
    e1.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / k], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


res, ms = {}, {}
granules = [int(g) for g in os.environ.get("GRANULES", "").split(",") if g] or [hp.LATTICE_GRANULE]
ms_by_granule = {}
for ex, gr in [("push", g) for g in granules] + [("allreduce", granules[0])]:
    it = sd.ShardedLatticeIteration(fid, gid, d, gen, tab, exchange=ex, granule=gr)
    it.step()
    it.step()
    torch.cuda.synchronize()
    out = (it.global_pool(), it.global_nfev())
    t = timed(it.step)
    it.close()
    if ex == "push":
        ms_by_granule[gr] = round(t, 4)
        if "push" in res and not (np.array_equal(res["push"][0], out[0]) and res["push"][1] == out[1]):
            res["push"] = (np.zeros(0), -1)        # granules disagree: reported as MISMATCH below
            continue
        if "push" not in res or t < ms["push"]:
            ms["push"] = t
        res.setdefault("push", out)
    else:
        res[ex], ms[ex] = out, t
ok = np.array_equal(res["push"][0], res["allreduce"][0]) and res["push"][1] == res["allreduce"][1]
if rank == 0:
    tmp = None

    def single():
        f, feas = hp.lattice_eval(fid, gid, d, gen, tab, 0, nv)
        return hp.compact_pool(hp.lattice_star(d, gen, f))
    idx, cnt = single()
    ref = idx[: int(cnt.item())].cpu().numpy()
    ok = ok and np.array_equal(ref, res["push"][0]) and res["push"][1] == nv
for_one = torch.zeros(1, device=dev)
if rank == 0:
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    for _ in range(3):
        single()
    e0.record()
    for _ in range(10):
        single()
    e1.record()
    e1.synchronize()
    for_one[0] = e0.elapsed_time(e1) / 10
    print(json.dumps({"check_sharded_lattice": "OK" if ok else "MISMATCH", "world": world, "d": d, "gen": gen,
                      "vertices": nv, "pool": int(len(ref)), "ms_one_gpu": round(float(for_one[0]), 4),
                      "ms_push": round(ms["push"], 4), "ms_push_by_granule": ms_by_granule, "ms_allreduce": round(ms["allreduce"], 4),
                      "speedup_push": round(float(for_one[0]) / ms["push"], 3)}))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
