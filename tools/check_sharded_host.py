"""Multi-GPU parity of the sharded HOST-buffer entry points (run under torchrun, one rank per GPU):
shgo_b200_ctx_shard_init / _connect / _iterate_sobol_host_shard_begin / _finish with numpy inputs against
the single-GPU shgo_b200_iterate_sobol_host on the whole problem.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29517 tools/check_sharded_host.py [log2n]
"""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from shgo_b200 import _lib, _sobol  # noqa: E402
from shgo_b200.hostshard import ShardedHostIteration  # noqa: E402

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, m, d = 1 << log2n, log2n, 6
rows = n // world
row0 = rank * rows
bounds = np.array([(-5.12, 5.12)] * d)
i = np.arange(row0, row0 + rows, dtype=np.int64)[:, None]
col = (i ^ (1 << np.arange(m, dtype=np.int64))[None, :]).astype(np.int32).reshape(-1)
rp = np.arange(row0 * m, (row0 + rows + 1) * m, m, dtype=np.int64)
it = ShardedHostIteration(local, world, rank, rows)
pools = []
for _ in range(3):            # three iterations: both f buffers are used
    pool, nfev = it.iterate(1, 0, d, bounds, rp, col, remote_at_ends=True, pool_cap=rows)
    pools.append(pool.copy())
it.close()
assert all(np.array_equal(pools[0], p) for p in pools)
parts = [None] * world
dist.all_gather_object(parts, (pools[0], nfev))
ok = True
if rank == 0:
    got = np.concatenate([p for p, _ in parts])
    nf = sum(v for _, v in parts)
    ii = np.arange(n, dtype=np.int64)[:, None]
    colw = (ii ^ (1 << np.arange(m, dtype=np.int64))[None, :]).astype(np.int32).reshape(-1)
    rpw = np.arange(0, (n + 1) * m, m, dtype=np.int64)
    ctx = ctypes.c_void_p()
    _lib.call("shgo_b200_ctx_create", local, ctypes.byref(ctx))
    ref = np.empty(n, np.int64)
    cnt, nfr = ctypes.c_uint64(), ctypes.c_uint64()
    sv = np.ascontiguousarray(_sobol.direction_numbers(d))
    lb, ub = bounds[:, 0].copy(), bounds[:, 1].copy()
    _lib.call("shgo_b200_iterate_sobol_host", ctx, 1, 0, d, 0, n, sv.ctypes.data, lb.ctypes.data, ub.ctypes.data,
              rpw.ctypes.data, colw.ctypes.data, ref.ctypes.data, n, ctypes.byref(cnt), ctypes.byref(nfr), None)
    _lib.lib().shgo_b200_ctx_destroy(ctx)
    ok = np.array_equal(got, ref[: cnt.value]) and nf == nfr.value
    print("check_sharded_host world=%d n=2^%d pool=%d nfev=%d : %s" % (world, log2n, len(got), nf, "OK" if ok else "MISMATCH"))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
