"""ncu driver for the GaussNet gather / scatter kernels on TEN DISTINCT 800x800 views per launch (P = 3): a working set of
0.8 GB of weights / indices / images per launch, far beyond the 126 MB L2, so the DRAM counters are a bandwidth figure and
the lts__t_sectors_op_red counters show what the 30.7 MB gradient table absorbs.  Three index patterns: clustered (+-400
around the pixel's own index: rendered geometry), the reference's typical 3x3 neighbourhood, uniformly random."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerfail_b200 import ops

dev = torch.device("cuda:0")
P, H, W, B = 3, 800, 800, 10
T = P * H * W
g = torch.Generator(device=dev).manual_seed(0)
table = torch.randn(P, H, W, 4, device=dev, generator=g) * 5
table[..., 3] = 255.0
for name in (sys.argv[1:] or ["clustered", "3x3", "random"]):
    base = torch.arange(H * W, device=dev).reshape(1, H, W, 1)
    if name == "3x3":
        offs = torch.tensor([0, 1, -1, W, -W, W + 1, -W - 1, W - 1], device=dev).reshape(1, 1, 1, 8)
        idx = (base + offs + torch.randint(0, P, (B, 1, 1, 1), device=dev, generator=g) * H * W).clamp_(0, T - 1).float()
    elif name == "clustered":
        idx = (base + torch.randint(0, P, (B, 1, 1, 1), device=dev, generator=g) * H * W
               + torch.randint(-400, 401, (B, H, W, 8), device=dev, generator=g)).clamp_(0, T - 1).float()
    else:
        idx = torch.randint(0, T, (B, H, W, 8), device=dev, generator=g).float()
    dist_ = torch.sort(torch.randn(B, H, W, 8, device=dev, generator=g).abs() * 0.01, dim=-1).values
    w_idx = ops.gauss_weights(torch.stack([dist_, idx], 1), 0.02)
    ori = torch.randint(0, 256, (B, H, W, 4), device=dev, generator=g, dtype=torch.uint8)
    gx = torch.randn(B, H, W, 4, device=dev, generator=g)
    grad = torch.zeros_like(table)
    del idx, dist_
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for it in range(3):
        e[0].record()
        x, x_rgba = ops.gauss_gather_fwd(table.reshape(-1, 4), w_idx, ori, 32.0)
        e[1].record()
        ops.gauss_scatter_bwd(None, gx, x, w_idx, ori, 32.0, table.shape, out=grad)
        e[2].record()
        torch.cuda.synchronize()
    f, b = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
    px = B * H * W
    print(f"{name}: {B} views per launch: gather {f * 1e3:.1f} us ({px * 228 / f / 1e6:.0f} GB/s algorithmic), "
          f"scatter {b * 1e3:.1f} us ({px * 228 / b / 1e6:.0f} GB/s algorithmic)")
