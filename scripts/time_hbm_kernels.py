"""Times the HBM-bound render kernels in isolation (CUDA events): compositing (coarse 64 / fine 192 samples), the
hierarchical resampling step, the compositing backward; prints achieved GB/s on the algorithmic bytes of SURVEY.md 8d."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerfail_b200 import ops
from oracle import synth

R = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
dev = torch.device("cuda:0")
K, _ = synth.intrinsics(800, 800)
rays = ops.get_ray_batch(800, 800, K, torch.tensor(synth.pose_spherical(30.0, -30.0, 4.0)[:3, :4]), 2.0, 6.0, device=dev)[:R].contiguous()
g = torch.Generator(device=dev).manual_seed(0)

def timed(fn, n=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

for S in (64, 192):
    z = ops.coarse_z(rays, S)
    raw = torch.randn(R, S, 4, device=dev, generator=g)
    ms = timed(lambda: ops.composite_fwd(raw, z, rays, None, True))
    by = R * (S * 24 + 36)
    print(f"composite_fwd  R={R} S={S:3d}: {ms * 1e3:8.1f} us  {by / ms / 1e6:7.1f} GB/s algorithmic ({by / 1e6:.0f} MB)")
z = ops.coarse_z(rays, 64)
w = torch.rand(R, 64, device=dev, generator=g)
zf, zs, zstd = ops.hierarchical(z, w, 128)
print(f"rays whose 128 new samples come out with an inversion: {float((zs[:, 1:] < zs[:, :-1]).any(1).float().mean()) * 100:.1f} %")
assert bool((zf[:, 1:] >= zf[:, :-1]).all())
ms = timed(lambda: ops.hierarchical(z, w, 128))
by = R * 1524
print(f"hierarchical   R={R} 64->192: {ms * 1e3:8.1f} us  {by / ms / 1e6:7.1f} GB/s algorithmic ({by / 1e6:.0f} MB)")
