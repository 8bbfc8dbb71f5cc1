"""Fused generate -> evaluate kernel alone: ms and fraction of the measured FP64 peak per functor/config.

    python tools/eval_probe.py [log2n]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from shgo_b200 import functors as F, hotpath as hp  # noqa: E402

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
only = sys.argv[2] if len(sys.argv) > 2 else None
peak = hp.fp64_peak_tflops()
print("measured FP64 peak (DFMA microbenchmark) %.2f TFLOP/s" % peak)
# flops per point (+,-,*,/,sqrt = 1, FMA = 2; transcendental calls not counted)
CASES = [
    ("styblinski_tang 8-D (cfg5)", F.F_STYBLINSKI_TANG, F.G_NONE, [(-5.0, 5.0)] * 8, lambda d: 2 * d + 7 * d + 1, log2n),
    ("rastrigin 6-D (cfg2)", F.F_RASTRIGIN, F.G_NONE, [(-5.12, 5.12)] * 6, lambda d: 2 * d + 5 * d + 1, min(log2n, 26)),
    ("ackley 10-D (cfg3 objective)", F.F_ACKLEY, F.G_NONE, [(-32.768, 32.768)] * 10, lambda d: 2 * d + 4 * d + 8, min(log2n, 26)),
    ("cattle-feed 4-D + g1,g2 (cfg4)", F.F_CATTLE_FEED, F.G_CATTLE_FEED, [(0.0, 1.0)] * 4, lambda d: 2 * d + 7 + 8 + 21, min(log2n, 26)),
    ("rosenbrock 5-D (cfg1 objective)", F.F_ROSENBROCK, F.G_NONE, [(0.0, 2.0)] * 5, lambda d: 2 * d + 8 * (d - 1), min(log2n, 26)),
    ("eggholder 2-D", F.F_EGGHOLDER, F.G_NONE, [(-512.0, 512.0)] * 2, lambda d: 2 * d + 14, min(log2n, 26)),
]
for name, fid, gid, bounds, flops, lg in CASES:
    if only and only not in name:
        continue
    n = 1 << lg
    tab = hp.SobolTables(bounds)
    f = torch.empty(n, dtype=torch.float64, device="cuda")
    feas = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        hp.eval_sobol(fid, gid, tab, 0, n, f=f, feasible=feas)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10):
        hp.eval_sobol(fid, gid, tab, 0, n, f=f, feasible=feas)
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / 10
    tf = flops(len(bounds)) * n / ms / 1e9
    print("%-34s n=2^%d  %8.3f ms  %7.2f Gpoints/s  %6.2f TFLOP/s algorithmic = %.3f of the FP64 peak  (%.0f GB/s written)"
          % (name, lg, ms, n / ms / 1e6, tf, tf / peak, 9 * n / ms / 1e6), flush=True)
