"""Star-kernel variants on the bench graph (index hypercube, random f): one subprocess per variant
because the switch SHGO_B200_STAR = "<bulk>,<warps per SM>,<max log2 lanes per survivor>" is read once.

    python tools/star_variants.py [log2n] [variant ...]      e.g.  27 1,0,5 0,0,5 1,24,4
"""
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def one(log2n):
    import torch
    sys.path.insert(0, ROOT)
    from shgo_b200 import hotpath as hp
    n, m = 1 << log2n, log2n
    dev = torch.device("cuda")
    row_ptr, col = hp.hypercube_csr(n, dev)
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    f = torch.rand(n, dtype=torch.float64, device=dev, generator=g)
    is_min = torch.empty(n, dtype=torch.uint8, device=dev)
    alg = 4 * col.numel() + 17 * n

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

    ms = timeit(lambda: hp.star_csr(row_ptr, col, f, is_min=is_min))
    pool = int(is_min.sum())
    fc = torch.zeros(n, dtype=torch.float64, device=dev)
    ms_c = timeit(lambda: hp.star_csr(row_ptr, col, fc, is_min=is_min))
    if os.environ.get("STAR_EXTRA"):
        from shgo_b200 import functors as F
        feas = torch.ones(n, dtype=torch.uint8, device=dev)
        nf = torch.zeros(1, dtype=torch.int64, device=dev)
        ms_nf = timeit(lambda: hp.star_csr(row_ptr, col, f, is_min=is_min, feasible=feas, nfev=nf))
        tab = hp.SobolTables([(-5.0, 5.0)] * 8)
        fst, _ = hp.eval_sobol(F.F_STYBLINSKI_TANG, 0, tab, 0, n)
        ms_st = timeit(lambda: hp.star_csr(row_ptr, col, fst, is_min=is_min))
        pool_st = int(is_min.sum())
        ms_st_nf = timeit(lambda: hp.star_csr(row_ptr, col, fst, is_min=is_min, feasible=feas, nfev=nf))
        print("   extra: random f + nfev %.3f ms | Styblinski-Tang of Sobol points %.3f ms (pool %d) | + nfev %.3f ms"
              % (ms_nf, ms_st, pool_st, ms_st_nf), flush=True)
    print("variant %-10s 2^%d  random f %7.3f ms %6.0f GB/s (%.3f of 6541)   constant f %7.3f ms %6.0f GB/s   pool %d"
          % (os.environ.get("SHGO_B200_STAR", "default"), log2n, ms, alg / ms / 1e6, alg / ms / 1e6 / 6541.1,
             ms_c, alg / ms_c / 1e6, pool), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        one(int(sys.argv[2]))
    else:
        log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 27
        variants = sys.argv[2:] or ["0,0,5", "1,0,5", "1,0,4", "1,0,3", "0,0,3", "1,24,5", "1,20,5"]
        for v in variants:
            env = dict(os.environ, SHGO_B200_STAR=v)
            subprocess.run([sys.executable, os.path.abspath(__file__), "--one", str(log2n)], env=env, check=False)
