"""Small stand-alone driver for ncu: a few launches of the fused MLP kernel (fine-pass shape) and of the HBM kernels."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerfail_b200 import ops
from oracle import synth

R = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
S = int(sys.argv[2]) if len(sys.argv) > 2 else 192
dev = torch.device("cuda:0")
m = ops.FusedMLP(device=dev)
m.update(synth.flat_params(synth.make_non_degenerate(synth.random_state_dict(1), 1)).to(dev))
K, _ = synth.intrinsics(800, 800)
rays = ops.get_ray_batch(800, 800, K, torch.tensor(synth.pose_spherical(30.0, -30.0, 4.0)[:3, :4]), 2.0, 6.0, device=dev)[:R].contiguous()
z = ops.coarse_z(rays, S)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(4):
    ev0.record()
    raw = m.forward_rays(rays, z)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    print(f"mlp R={R} S={S}: {ms:.3f} ms  {R * S * 1186816 / ms / 1e9:.1f} TFLOP/s")
m.status()
out = ops.composite_fwd(raw, z, rays, None, True)
zf = ops.hierarchical(z[:, :64].contiguous(), out[3][:, :64].contiguous(), 128)
torch.cuda.synchronize()
