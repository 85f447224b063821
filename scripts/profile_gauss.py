"""Driver for ncu / timing of the GaussNet gather + scatter kernels at 800x800, P=3 (BASELINE config 3 shapes)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerfail_b200 import ops

dev = torch.device("cuda:0")
P, H, W = 3, 800, 800
T = P * H * W
g = torch.Generator(device=dev).manual_seed(0)
table = torch.randn(P, H, W, 4, device=dev, generator=g) * 5
for locality in (0, 400, T):
    base = torch.arange(H * W, device=dev).reshape(1, H, W, 1)
    if locality == 0:      # realistic: a pixel's 8 neighbours are the points of its own 3x3 neighbourhood -> heavy sharing
        offs = torch.tensor([0, 1, -1, W, -W, W + 1, -W - 1, W - 1], device=dev).reshape(1, 1, 1, 8)
        idx = (base + offs).clamp_(0, T - 1).float()
    elif locality < T:
        idx = (base + torch.randint(0, P, (1,), device=dev, generator=g) * H * W + torch.randint(-locality, locality + 1, (1, H, W, 8), device=dev, generator=g)).clamp_(0, T - 1).float()
    else:
        idx = torch.randint(0, T, (1, H, W, 8), device=dev, generator=g).float()
    dist_ = torch.sort(torch.randn(1, H, W, 8, device=dev, generator=g).abs() * 0.01, dim=-1).values
    w_idx = ops.gauss_weights(torch.stack([dist_, idx], 1), 0.02)
    ori = torch.randint(0, 256, (1, H, W, 4), device=dev, generator=g, dtype=torch.uint8)
    gx = torch.randn(1, H, W, 4, device=dev, generator=g)
    grad = torch.zeros_like(table)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for it in range(5):
        e[0].record()
        x, x_rgba = ops.gauss_gather_fwd(table.reshape(-1, 4), w_idx, ori, 32.0)
        e[1].record()
        ops.gauss_scatter_bwd(None, gx, x, w_idx, ori, 32.0, table.shape, out=grad)
        e[2].record()
        torch.cuda.synchronize()
    f, b = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
    px = H * W
    print(f"locality +-{locality}: gather {f * 1e3:.1f} us ({px * 228 / f / 1e6:.0f} GB/s algorithmic), scatter {b * 1e3:.1f} us ({px * 228 / b / 1e6:.0f} GB/s algorithmic)")
