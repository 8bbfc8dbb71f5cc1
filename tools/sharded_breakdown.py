"""Per-stage timing of the sharded (p2p) iteration, run under torchrun; rank 0 prints the table."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from shgo_b200 import dist as sd, functors as F, hotpath as hp
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 26
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n, m, d = 1 << log2n, log2n, 8
tab = hp.SobolTables([(-5.0, 5.0)] * d, dev)
row0, rows = sd.shard_range(n, world, rank)
rp_l = torch.arange(row0 * m, (row0 + rows + 1) * m, m, dtype=torch.int64, device=dev)
col_l = torch.empty((rows, m), dtype=torch.int32, device=dev)
bits = (1 << torch.arange(m, dtype=torch.int32, device=dev))[None, :]
for s in range(0, rows, 1 << 22):
    e = min(rows, s + (1 << 22))
    torch.bitwise_xor(torch.arange(row0 + s, row0 + e, dtype=torch.int32, device=dev)[:, None], bits, out=col_l[s:e])
col_l = col_l.reshape(-1)
it = sd.ShardedSobolIteration(F.F_STYBLINSKI_TANG, 0, tab, n, rp_l, col_l, exchange="p2p", pool_cap=rows // 8,
                              remote_at_ends=True)
for _ in range(5): it.step()
torch.cuda.synchronize(); dist.barrier()
names = ["eval", "local star", "barrier wait", "candidates+remote", "pool compaction"]
acc = [0.0] * 5
K = 20
for _ in range(K):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    it.peer = it.peers[it._k & 1]; it.f_my = it.peer.local; it._k += 1
    ev[0].record()
    hp.eval_sobol(it.functor, it.gset, it.tab, it.row0, it.rows, f=it.f_my, feasible=it.feas)
    ev[1].record()
    it._barrier_async()
    hp.star_csr_local(it.row_ptr, it.col, it.row0, it.f_my, it.feas, it.is_min, it.nfev, it.off0, cand_ws=it._cand_ws, remote_at_ends=it.remote_at_ends)
    ev[2].record()
    it._barrier_wait()
    ev[3].record()
    hp.star_csr_remote(it.row_ptr, it.col, it.row0, it.peer.table, it.peer.own_ptr, it.world, it.rows, it.rank,
                       it.is_min, it._remote_tmp, it.off0, remote_at_ends=True, cand_ws=it._cand_ws)
    ev[4].record()
    hp.compact_pool(it.is_min, base=it.row0, work=it.work)
    ev[5].record()
    torch.cuda.synchronize()
    for i in range(5): acc[i] += ev[i].elapsed_time(ev[i + 1])
cand = int((it.is_min == 2).sum().item())
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(K): it.step()
e1.record(); torch.cuda.synchronize()
if rank == 0:
    print("world %d n=2^%d rows/gpu %d  pipelined step %.4f ms" % (world, log2n, rows, e0.elapsed_time(e1) / K))
    for nm, a in zip(names, acc): print("  %-20s %.4f ms" % (nm, a / K))
    print("  total %.4f ms" % (sum(acc) / K))
it.close(); dist.barrier(); dist.destroy_process_group()
