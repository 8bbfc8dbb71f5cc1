#!/usr/bin/env python
"""bench.py -- sample points evaluated + classified per second (BASELINE.json's metric).

Workload (config 5 of BASELINE.json, the one the metric's 1/2/4/8-GPU sweep is quoted on):
Styblinski-Tang 8-D on [-5,5]^8, unscrambled Sobol indices [0, n), n = 2^28 by default, classified
on the SYNTHETIC index-hypercube CSR of BASELINE.md section 5 (vertex i ~ i XOR 2^j, degree
log2 n; explicit int64 row_ptr / int32 col).  THIS IS NOT A DELAUNAY GRAPH: qhull cannot produce
one at this size; parity on real qhull graphs is covered by tests/.

One step = one shgo iteration of the hot path: fused generate->scale->evaluate->mask (K1+K2),
[exchange of f between ranks], CSR star test (K4), pool compaction, nfev count.
Strong scaling: n is fixed, rank r owns Sobol indices / CSR rows [r n/N, (r+1) n/N).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--log2n 28] [--impl reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D = 8
BOUNDS = [(-5.0, 5.0)] * D
FUNCTOR = "styblinski_tang"
FLOPS_PER_POINT = 2 * D + 7 * D + 1      # scaling 2d + functor 7d+1
KERNELS_PER_STEP = 5                      # eval, star(+nfev), 3x compaction (+1 at N>1: the phase-2 kernel)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2n", type=int, default=28, help="log2 of the total number of points")
    ap.add_argument("--e2e-log2n", type=int, default=26,
                    help="log2 n of the host-buffer (e2e) leg AND of the CPU arm's sample (same problem in both)")
    ap.add_argument("--cpu-log2n", type=int, default=None, help="override the CPU sample size (default: --e2e-log2n)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "fused", "allgather"],
                    help="how neighbour f values cross GPUs (N > 1)")
    ap.add_argument("--no-filter", action="store_true", help="p2p exchange without the block-minimum filter")
    ap.add_argument("--chunks", type=int, default=1, help="p2p exchange: classify the shard in chunks (overlap)")
    ap.add_argument("--phase1-warps", type=int, default=0, help="p2p + chunks: resident warps per SM of phase 1")
    ap.add_argument("--workload", default="cfg5", choices=["cfg5", "delaunay"],
                    help="cfg5: the headline sweep (synthetic CSR); delaunay: a real qhull mesh of Sobol points "
                         "(config 2 in 2-D at n = 2^20 by default), star test in qhull's and in Z-order numbering")
    ap.add_argument("--delaunay-dim", type=int, default=2)
    ap.add_argument("--delaunay-log2n", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle port (the reference itself is per-vertex Python at ~1.5e4 points/s and
# cannot build this graph) on all host threads, bounded sample.
# ----------------------------------------------------------------------------------------------
def cpu_run(log2n, repeats):
    import numpy as np
    from oracle import oracle as orc
    orc.set_threads(os.cpu_count() or 1)     # all host threads, also under torchrun (OMP_NUM_THREADS=1)
    n = 1 << log2n
    b = np.array(BOUNDS)
    sv = orc.sobol_dirnums(D)
    rp, col = orc.hypercube_csr(n)
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        f, feas = orc.sobol_eval(sv, FUNCTOR, None, 0, n, b[:, 0], b[:, 1])
        is_min = orc.star_csr(rp, col, f)
        pool = orc.compact_pool(is_min)
        nfev = orc.nfev(rp, feas)
        times.append(time.perf_counter() - t0)
    return n, times, len(pool), nfev, orc.num_threads()


def python_driver_run(n=4096, d=3):
    """The reference's own Python path (scipy's vendored copy of the driver + Complex; the GPU box
    has no /root/reference) on a bounded sample: Sobol n points in d-D -> qhull -> vf_to_vv ->
    process_pools -> minimizers, one core.  Returns points/s with and without the qhull time."""
    import numpy as np
    from scipy.optimize import _shgo as drv
    from scipy.spatial import Delaunay

    def st(x):
        return 0.5 * np.sum(x ** 4 - 16 * x ** 2 + 5 * x)

    s = drv.SHGO(st, [(-5.0, 5.0)] * d, n=n, iters=1, sampling_method="sobol")
    t0 = time.perf_counter()
    s.iterate_complex()
    s.minimizers()
    total = time.perf_counter() - t0
    t0 = time.perf_counter()
    Delaunay(s.C)
    qhull = time.perf_counter() - t0
    return {"value": n / total, "value_without_qhull": n / max(total - qhull, 1e-9), "unit": "points/s",
            "cores": 1, "sample": "scipy.optimize._shgo (vendored copy of the reference's driver and "
            "Python Complex): Styblinski-Tang %d-D, Sobol n=%d, iterate_complex + minimizers; "
            "qhull %.2f s of %.2f s" % (d, n, qhull, total)}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, times, npool, nfev, cores = cpu_run(args.cpu_log2n, args.warmup + args.steps)
    t = times[args.warmup:]
    ms = 1e3 * sum(t) / len(t)
    val = n / (ms * 1e-3)
    sample = ("each step = the same hot path (generate, evaluate, mask, star test, pool) run by the "
              "oracle C port (oracle/shgo_oracle.c, OpenMP, all host threads) on a bounded sample of "
              "2^%d points with its degree-%d hypercube graph; the Python reference itself runs "
              "~1.5e4 points/s single-core (BASELINE.md)" % (args.cpu_log2n, args.cpu_log2n))
    print(json.dumps({
        "impl": "reference", "metric": "sample points evaluated+classified/sec", "value": val,
        "unit": "points/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": config_of(args),
        "sample_points_per_step": n,
        "cpu_baseline": {"value": val, "unit": "points/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": val, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


EXCHANGES = {
    "fused": "none: f shards peer-mapped over NVLink (CUDA IPC); the star kernel reads its candidates' remote "
             "neighbours in place, inline, once the owner has published its shard; 1 stream-ordered barrier per step",
    "p2p": "none: f shards peer-mapped over NVLink (CUDA IPC), the candidates' remote neighbour values read in place "
           "by a second kernel, pre-filtered by group minima pushed in bulk; 1 stream-ordered barrier per step",
    "allgather": "nccl all_gather of f", "none": "none"}


def config_of(args):
    """Identical in both arms by construction (everything comes from the command line): the workload, the
    size both the CPU arm and the e2e leg run, and how the GPU arm treats L2 and the exchange."""
    n, world = 1 << args.log2n, max(args.gpus, 1)
    return {"workload": workload_name(args.log2n), "points": n,
            "e2e_and_cpu_points": 1 << min(args.e2e_log2n, args.log2n),
            "note": "value: device-resident 2^%d points; e2e (host buffers through the C ABI) and the CPU arm both "
                    "run the same problem at 2^%d points (degree-%d hypercube graph): the host CSR of the full "
                    "size is 32 GB" % (args.log2n, min(args.e2e_log2n, args.log2n), min(args.e2e_log2n, args.log2n)),
            "rows_per_gpu": n // world,
            "l2": "inputs (col = %.1f GB per GPU) far exceed the 126 MB L2; no flush needed"
                  % (n // world * args.log2n * 4 / 1e9),
            "exchange": EXCHANGES[args.exchange if world > 1 else "none"]}


def workload_name(log2n):
    return ("cfg5: Styblinski-Tang 8-D [-5,5]^8, Sobol n=2^%d, synthetic index-hypercube CSR "
            "(degree %d, not Delaunay)" % (log2n, log2n))


# ----------------------------------------------------------------------------------------------
def pool_digest(local_pool, row0, rows, n, world, dist, dev):
    """N-invariant fingerprint of the global minimiser pool: sha256 over the digests of the pool
    restricted to fixed index blocks of 2^22 rows (a rank owns whole blocks), gathered in rank order."""
    import hashlib
    import numpy as np
    import torch
    blk = min(1 << 22, rows)
    assert rows % blk == 0
    edges = np.arange(row0, row0 + rows + 1, blk)
    cut = np.searchsorted(local_pool, edges)
    mine = b"".join(hashlib.sha256(np.ascontiguousarray(local_pool[cut[i]:cut[i + 1]]).tobytes()).digest()
                    for i in range(len(edges) - 1))
    t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(dev)
    if world > 1:
        out = torch.empty(world * t.numel(), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(out, t)
        t = out
    return hashlib.sha256(t.cpu().numpy().tobytes()).hexdigest()[:16]


def delaunay_workload(args):
    """A REAL Delaunay graph (scipy.spatial.Delaunay = qhull, on the host, outside every timed region) of
    Sobol points: K3 (vf_to_vv -> CSR), the Z-order renumbering, and K4 (star test) in both numberings,
    each launch timed alone with CUDA events after an L2 flush (the whole problem fits the 126 MB L2)."""
    import numpy as np
    import torch
    from scipy.spatial import Delaunay
    from shgo_b200 import functors as F
    from shgo_b200 import hotpath as hp
    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback"
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    d, n = args.delaunay_dim, 1 << args.delaunay_log2n
    bounds = [(-5.12, 5.12)] * d
    tab = hp.SobolTables(bounds, dev)
    pts = hp.sobol_points(tab, 0, n, "aos").cpu().numpy()
    t0 = time.perf_counter()
    simp = Delaunay(pts).simplices.astype(np.int32)
    qhull_s = time.perf_counter() - t0
    S = simp.shape[0]
    simp_d = torch.from_numpy(simp).to(dev)
    x_soa = hp.sobol_points(tab, 0, n, "soa")
    flush = torch.empty(1 << 28, dtype=torch.uint8, device=dev)      # 256 MB > L2
    hbm_peak, peak_src = peaks()

    def timed(fn, k):
        """mean ms of k launches, each after an L2 flush, timed alone"""
        out, ms = None, []
        for _ in range(k + 2):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            e1.record()
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
        return out, sum(ms[2:]) / k

    k = max(args.steps, 5)
    (rp, col), k3_ms = timed(lambda: hp.csr_from_simplices(simp_d, n), k)
    (order, newid), mort_ms = timed(lambda: hp.morton_order(x_soa, bounds), k)
    s3, rel_ms = timed(lambda: hp.relabel3(simp_d, n, newid), k)
    (rp_m, col_m), k3m_ms = timed(lambda: hp.csr_from_simplices(s3, n), k)
    nnz = int(col.numel())
    (f, feas), eval_ms = timed(lambda: hp.eval_sobol(F.F_RASTRIGIN, F.G_NONE, tab, 0, n), k)
    flags = torch.empty(n, dtype=torch.uint8, device=dev)
    flags_m = torch.empty(n, dtype=torch.uint8, device=dev)
    back = torch.empty(n, dtype=torch.uint8, device=dev)
    f_m = torch.empty(n, dtype=torch.float64, device=dev)
    _, k4_ms = timed(lambda: hp.star_csr(rp, col, f, is_min=flags), k)
    _, gat_ms = timed(lambda: hp.gather(f, order, out=f_m), k)
    _, k4m_ms = timed(lambda: hp.star_csr(rp_m, col_m, f_m, is_min=flags_m), k)
    _, sca_ms = timed(lambda: hp.scatter_u8(flags_m, order, back), k)
    work = hp.PoolWorkspace(n, n, dev)
    _, pool_ms = timed(lambda: hp.compact_pool(back, work=work), k)
    pool = work.idx[: int(work.count[0].item())].cpu().numpy()
    same = bool(torch.equal(back, flags))
    alg4 = 4 * nnz + 17 * n
    alg3 = 12 * S + 8 * (n + 1) + 4 * nnz
    step_ms = eval_ms + gat_ms + k4m_ms + sca_ms + pool_ms

    def roof(alg, ms):
        a = alg / (ms * 1e-3) / 1e9
        return {"achieved": a, "frac": a / hbm_peak, "ms_per_launch": ms, "algorithmic_bytes_per_launch": alg}

    out = {
        "metric": "sample points evaluated+classified/sec", "value": n / (step_ms * 1e-3), "unit": "points/s",
        "n_gpus": 1, "steps": k, "warmup": 2, "ms_per_step": step_ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "delaunay: Rastrigin %d-D [-5.12,5.12]^%d, Sobol n=2^%d, REAL Delaunay graph "
                               "(scipy.spatial.Delaunay / qhull on the host, %.1f s, outside the timed regions): "
                               "%d simplices, %d directed edges" % (d, d, args.delaunay_log2n, qhull_s, S, nnz),
                   "points": n, "l2": "the problem fits the 126 MB L2: every launch is timed alone after a 256 MB "
                                      "flush write"},
        "roofline": dict(roof(alg4, k4m_ms), bound="hbm", kernel="star_kernel (csrc/star.cu), Z-order numbering",
                         peak=hbm_peak, unit="GB/s", traffic=None, peak_source=peak_src),
        "k4_qhull_numbering": roof(alg4, k4_ms),
        "k3_csr_from_simplices": dict(roof(alg3, k3_ms), note="12 S read + 8 (n+1) + 4 nnz written; hash-set "
                                      "de-duplication and the row sort are extra traffic"),
        "k3_in_z_order": roof(alg3, k3m_ms),
        "stages_ms": {"eval": eval_ms, "gather_f": gat_ms, "star_z_order": k4m_ms, "scatter_flags": sca_ms,
                      "pool": pool_ms, "once_per_mesh": {"morton_order": mort_ms, "relabel": rel_ms,
                                                         "csr_z_order": k3m_ms, "csr_qhull_order": k3_ms}},
        "pool_size": int(len(pool)), "pool_identical_in_both_numberings": same,
        "gpu_launches": 7 * k,
    }
    if not args.no_cpu:
        from oracle import oracle as orc
        orc.set_threads(os.cpu_count() or 1)
        t0 = time.perf_counter()
        rp_o, col_o = orc.vf_to_vv(simp, n)
        t1 = time.perf_counter()
        fo, _ = orc.eval_points("rastrigin", None, pts)
        pool_o = orc.compact_pool(orc.star_csr(rp_o, col_o, f.cpu().numpy()))
        t2 = time.perf_counter()
        out["cpu_baseline"] = {"value": n / (t2 - t1), "unit": "points/s", "cores": orc.num_threads(), "kind": "port",
                               "sample": "oracle C port: evaluate + star test + pool on the same mesh (%.3f s); "
                                         "vf_to_vv %.2f s; qhull %.1f s" % (t2 - t1, t1 - t0, qhull_s)}
        out["pool_equals_oracle"] = bool(np.array_equal(pool, pool_o)) and bool(
            np.array_equal(rp.cpu().numpy(), rp_o) and np.array_equal(col.cpu().numpy(), col_o))
    print(json.dumps(out))


def main():
    args = parse()
    if args.workload == "delaunay" and args.impl != "reference":
        return delaunay_workload(args)
    if args.cpu_log2n is None:
        args.cpu_log2n = min(args.e2e_log2n, args.log2n)
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    from shgo_b200 import functors as F
    from shgo_b200 import hotpath as hp
    from shgo_b200 import _lib
    import ctypes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs CUDA devices; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d" % args.gpus

    n = 1 << args.log2n
    m = args.log2n
    assert n % world == 0
    rows = n // world
    row0 = rank * rows
    fid = F.F_STYBLINSKI_TANG
    tab = hp.SobolTables(BOUNDS, dev)

    # ---- graph shard (inputs resident in HBM before the timed region) -----------------------
    from shgo_b200 import dist as sdist
    row_ptr_l = torch.arange(row0 * m, (row0 + rows + 1) * m, m, dtype=torch.int64, device=dev)
    col_l = torch.empty((rows, m), dtype=torch.int32, device=dev)
    bits = (1 << torch.arange(m, dtype=torch.int32, device=dev))[None, :]
    for s in range(0, rows, 1 << 22):
        e = min(rows, s + (1 << 22))
        i = torch.arange(row0 + s, row0 + e, dtype=torch.int32, device=dev)[:, None]
        torch.bitwise_xor(i, bits, out=col_l[s:e])
    col_l = col_l.reshape(-1)
    del bits, i
    # rows list i ^ 2^j in ascending j, so the neighbours owned by other GPUs (the top log2 N bits)
    # are the last entries of every row: the remote_at_ends promise of the sharded star test holds
    it = sdist.ShardedSobolIteration(fid, F.G_NONE, tab, n, row_ptr_l, col_l, exchange=args.exchange,
                                     pool_cap=rows // 8, remote_at_ends=True, chunks=args.chunks,
                                     phase1_warps=args.phase1_warps, filter_remote=not args.no_filter)
    stream = torch.cuda.current_stream()

    def step(ev=None):
        it.step(ev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    # the timed region is tens of ms; keep the same load running (all ranks, fixed count) long
    # enough for nvidia-smi (100 ms period, slow to start) to sample it, so that the clocks line
    # describes this workload
    for _ in range(40):
        step()
    barrier()
    # Stage breakdown from events recorded INSIDE the timed region (start / after eval / after star
    # of every step): eval + star + pool == step by construction.  At N > 1 the star stage of a rank
    # includes its wait for the slowest peer's evaluation (the barrier), which is on that rank's
    # critical path; the per-stage figures are averaged over ranks, the step time is the max.
    stream = torch.cuda.current_stream()
    marks = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    # one more untimed step right in front of the timed region, in stream order: its device-side barrier
    # lines the ranks' streams up to within a few microseconds, where dist.barrier() + synchronize leave
    # them up to ~0.5 ms apart (that skew would be charged to the first timed step of the early ranks)
    step()
    t0.record(stream)
    for k in range(args.steps):
        marks[k][0].record(stream)
        step(marks[k][1:])
    t1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = t0.elapsed_time(t1)
    ends = [marks[k + 1][0] for k in range(args.steps - 1)] + [t1]
    eval_in = sum(m[0].elapsed_time(m[1]) for m in marks) / args.steps
    star_ms = sum(m[1].elapsed_time(m[2]) for m in marks) / args.steps
    pool_ms = sum(m[2].elapsed_time(e) for m, e in zip(marks, ends)) / args.steps
    tt = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    stages = torch.tensor([eval_in, star_ms, pool_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(stages)
        stages /= world
    ms_step = float(tt[0].item()) / args.steps
    eval_in, star_ms, pool_ms = (float(x) for x in stages.tolist())
    value = n / (ms_step * 1e-3)
    local_pool = it.local_pool().cpu().numpy()
    pool_sha = pool_digest(local_pool, row0, rows, n, world, dist, dev)
    pc = torch.tensor([int(it.work.count[0].item()), int(it.nfev.item())], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(pc)

    # eval kernel alone (for the FP64 side of the roofline)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(5):
        hp.eval_sobol(fid, F.G_NONE, tab, row0, rows, f=it.f_my, feasible=it.feas)
    e1.record(stream)
    torch.cuda.synchronize()
    eval_ms = e0.elapsed_time(e1) / 5

    # ---- e2e: HOST buffers through the C ABI, copies inside the timed region ------------------
    e2e = None
    if not args.no_e2e:
        ne = 1 << min(args.e2e_log2n, args.log2n)
        me = min(args.e2e_log2n, args.log2n)
        re_ = ne // world
        r0e = rank * re_
        it.close()
        del col_l, row_ptr_l
        torch.cuda.empty_cache()
        rp_h = torch.arange(0, (re_ + 1) * me, me, dtype=torch.int64).pin_memory()
        col_d = torch.empty((re_, me), dtype=torch.int32, device=dev)
        bits = (1 << torch.arange(me, dtype=torch.int32, device=dev))[None, :]
        for s in range(0, re_, 1 << 22):
            e = min(re_, s + (1 << 22))
            i = torch.arange(r0e + s, r0e + e, dtype=torch.int32, device=dev)[:, None]
            torch.bitwise_xor(i, bits, out=col_d[s:e])
        col_h = torch.empty(re_ * me, dtype=torch.int32).pin_memory()
        col_h.copy_(col_d.reshape(-1))
        del col_d
        torch.cuda.empty_cache()
        pool_h = torch.empty(max(re_ // 8, 1), dtype=torch.int64).pin_memory()
        cnt, nf = ctypes.c_uint64(), ctypes.c_uint64()
        if world == 1:
            ctx = ctypes.c_void_p()
            _lib.call("shgo_b200_ctx_create", local, ctypes.byref(ctx))
            sv_h, lb_h, ub_h = tab.sv_host, tab.lb_host, tab.ub_host

            def e2e_step():
                _lib.call("shgo_b200_iterate_sobol_host", ctx, fid, F.G_NONE, D, 0, ne, sv_h.ctypes.data,
                          lb_h.ctypes.data, ub_h.ctypes.data, rp_h.data_ptr(), col_h.data_ptr(),
                          pool_h.data_ptr(), pool_h.numel(), ctypes.byref(cnt), ctypes.byref(nf), None)
        else:
            # multi-GPU: the sharded C-ABI entry points (one ctx per rank, same peer-memory data path as
            # `value`): host CSR shard uploaded over the rank's own PCIe link, pool indices read back
            from shgo_b200.hostshard import ShardedHostIteration
            shard = ShardedHostIteration(local, world, rank, re_)
            # global offsets / ids, as the ABI wants them
            rp_h = torch.arange(r0e * me, (r0e + re_ + 1) * me, me, dtype=torch.int64).pin_memory()

            def e2e_step():
                shard.iterate(fid, F.G_NONE, D, BOUNDS, rp_h, col_h, remote_at_ends=True,
                              pool_cap=pool_h.numel(), pool_out=pool_h)
                cnt.value = shard.pool_count

        for _ in range(2):
            e2e_step()
        barrier()
        w0 = time.perf_counter()
        ksteps = max(3, min(args.steps, 5))
        for _ in range(ksteps):
            e2e_step()
        barrier()
        wall = (time.perf_counter() - w0) / ksteps
        wt = torch.tensor([wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(wt, op=dist.ReduceOp.MAX)
        wall = float(wt.item())
        h2d = (re_ + 1) * 8 + re_ * me * 4 + D * 30 * 4 + 2 * D * 8
        d2h = 16 + int(cnt.value) * 8
        e2e = {"value": ne / wall, "unit": "points/s", "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": d2h * world, "points": ne, "ms_per_step": wall * 1e3,
               "note": "host CSR (pinned) uploaded, pool indices read back, every step; "
                       + ("C-ABI shgo_b200_iterate_sobol_host" if world == 1 else
                          "C-ABI shgo_b200_iterate_sobol_host_shard_begin / _finish, one ctx per rank: every "
                          "rank uploads its CSR shard over its own PCIe link, f stays in peer-mapped shards")}
        if world == 1:
            _lib.lib().shgo_b200_ctx_destroy(ctx)
        else:
            shard.close()

    it.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_peak, peak_src = peaks()
    # dram bytes of ONE launch of the dominant kernel from a committed `ncu --set full` capture; only
    # quoted while the kernel source is the one that was profiled (sha of csrc/star.cu in the file)
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "r2_star_traffic.json")
    if os.path.exists(tp):
        import hashlib
        with open(tp) as fh:
            tj = json.load(fh)
        with open(os.path.join(ROOT, "shgo_b200", "csrc", "star.cu"), "rb") as fh:
            sha = hashlib.sha256(fh.read()).hexdigest()[:16]
        if tj.get("log2n") == args.log2n and tj.get("n_gpus") == world and tj.get("star_cu_sha16") == sha:
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
            traffic_src = tj["source"]
        elif tj.get("log2n") == args.log2n and tj.get("n_gpus") == world:
            traffic_src = "stale: csrc/star.cu changed since " + tj["source"]
    nnz_l = rows * m
    alg_bytes = 4 * nnz_l + 17 * rows
    achieved = alg_bytes / (star_ms * 1e-3) / 1e9
    fp64_peak = hp.fp64_peak_tflops()
    eval_tflops = FLOPS_PER_POINT * rows / (eval_ms * 1e-3) / 1e12
    out = {
        "metric": "sample points evaluated+classified/sec", "value": value, "unit": "points/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": config_of(args),
        "roofline": {"bound": "hbm", "kernel": "star_kernel (csrc/star.cu)", "achieved": achieved,
                     "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": star_ms},
        "eval_kernel": {"ms_per_launch": eval_ms, "tflops_fp64_algorithmic": eval_tflops,
                        "fp64_peak_measured_tflops": fp64_peak,
                        "frac_fp64": eval_tflops / fp64_peak if fp64_peak else None,
                        "hbm_write_gbs": 9 * rows / (eval_ms * 1e-3) / 1e9},
        "stages_ms": {"eval": eval_in, "star": star_ms, "pool": pool_ms,
                      "note": "CUDA events inside the timed region, mean over steps and ranks; they add up "
                              "to the mean step; at N > 1 the star stage includes the wait for the slowest "
                              "peer's evaluation"},
        "pool_size": int(pc[0].item()), "nfev": int(pc[1].item()), "pool_sha16": pool_sha,
        "gpu_launches": (KERNELS_PER_STEP + {"fused": 2, "p2p": 1}.get(it.exchange, 0)) * args.steps,
        "clocks": clocks,
    }
    if e2e:
        out["e2e"] = e2e
    if not args.no_cpu and world == 1:      # reported at N = 1 only
        ncpu, times, npool, nfev_c, cores = cpu_run(args.cpu_log2n, 3)
        best = min(times)
        out["cpu_baseline"] = {
            "value": ncpu / best, "unit": "points/s", "cores": cores, "kind": "port",
            "sample": "oracle C port (OpenMP, %d threads) of the same step on 2^%d points (degree "
                      "%d graph), best of 3; the Python reference itself runs ~1.5e4 points/s "
                      "single-core (BASELINE.md section 2)" % (cores, args.cpu_log2n, args.cpu_log2n)}
        try:
            out["cpu_baseline"]["python_reference"] = python_driver_run()
        except Exception as e:          # informational only
            out["cpu_baseline"]["python_reference"] = {"unavailable": repr(e)[:120]}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
