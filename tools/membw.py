"""Read-only / copy bandwidth reference on this GPU (torch kernels), to calibrate the HBM ceiling."""
import torch
x = torch.empty(1 << 30, dtype=torch.float64, device="cuda").normal_()   # 8 GiB
y = torch.empty_like(x)
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    best = 1e9
    for _ in range(n):
        e0.record(); fn(); e1.record(); e1.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: x.sum()); print("read  (sum)  %.1f GB/s" % (x.numel()*8/ms/1e6))
ms = t(lambda: x.max()); print("read  (max)  %.1f GB/s" % (x.numel()*8/ms/1e6))
ms = t(lambda: y.copy_(x)); print("copy         %.1f GB/s (r+w)" % (2*x.numel()*8/ms/1e6))
ms = t(lambda: y.fill_(1.0)); print("write (fill) %.1f GB/s" % (x.numel()*8/ms/1e6))
