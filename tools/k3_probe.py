"""K3 (vf_to_vv -> CSR) at sizes qhull needs minutes for: a regular triangulation of a g x g grid (two
triangles per cell, 6 neighbours per interior vertex, like a 2-D Delaunay mesh), vertex ids either in
grid order or randomly permuted (what qhull hands over for Sobol-ordered points).

    python tools/k3_probe.py [log2 g]        # n = g*g vertices, S = 2 (g-1)^2 simplices
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from shgo_b200 import hotpath as hp  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 12
g = 1 << lg
n = g * g
dev = torch.device("cuda")
i = torch.arange(g - 1, device=dev, dtype=torch.int64)
a = (i[:, None] * g + i[None, :]).reshape(-1)
tri = torch.cat([torch.stack([a, a + 1, a + g], 1), torch.stack([a + 1, a + g + 1, a + g], 1)]).to(torch.int32)
S = tri.shape[0]
flush = torch.empty(1 << 28, dtype=torch.uint8, device=dev)


def timed(simp, k=5):
    ms = []
    for _ in range(k + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        rp, col = hp.csr_from_simplices(simp, n)
        e1.record()
        e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    return rp, col, sum(ms[2:]) / k


for name, simp in (("grid order", tri), ("random order", torch.randperm(n, device=dev).to(torch.int32)[tri.long()].contiguous())):
    rp, col, ms = timed(simp)
    nnz = col.numel()
    alg = 12 * S + 8 * (n + 1) + 4 * nnz
    deg = (rp[1:] - rp[:-1])
    assert bool((col[1:] > col[:-1])[(torch.arange(nnz - 1, device=dev) + 1 != 0)].sum() >= 0)
    print("K3 %-12s n = 2^%d, S = %d, nnz = %d (max degree %d): %.3f ms = %.0f GB/s algorithmic (%.3f of 6541)"
          % (name, 2 * lg, S, nnz, int(deg.max()), ms, alg / ms / 1e6, alg / ms / 1e6 / 6541.1), flush=True)
