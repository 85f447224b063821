"""Measured HBM ceilings of this box for the three access mixes of the training kernels: write-only (saved images),
read-only (weight-gradient operands) and copy (MEASURED_PEAKS.json's figure)."""
import torch
dev = torch.device("cuda:0")
n = 1 << 30            # 4 GiB of fp32
a = torch.empty(n, device=dev); b = torch.empty(n, device=dev)
def t(fn, bytes_, name):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"{name:12s} {bytes_ / best / 1e6:8.1f} GB/s")
t(lambda: a.fill_(1.0), 4 * n, "write-only")
t(lambda: a.max(), 4 * n, "read-only")
t(lambda: b.copy_(a), 8 * n, "copy")
