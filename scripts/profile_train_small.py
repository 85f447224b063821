"""One rank's share of a strong-scaled retraining step (512 of 4096 rays) as a CUDA graph with PeerAdam, world size 1:
what an 8-GPU data-parallel step costs per GPU apart from the peer traffic.  python scripts/profile_train_small.py [rays] [iters]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerfail_b200 as nb
from nerfail_b200 import dist as nd, ops, train as ntrain
from oracle import synth
from bench import LegoArgs, H, W

n_rays = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda:0")
_, kw, *_ = nb.create_nerf(LegoArgs(), device=dev)
kw["network_fn"].load_state_dict(synth.make_non_degenerate(synth.random_state_dict(0), 0))
kw["network_fine"].load_state_dict(synth.make_non_degenerate(synth.random_state_dict(1), 1))
K, _ = synth.intrinsics(H, W)
rays_all = ops.get_ray_batch(H, W, K, torch.tensor(synth.camera_ring(8)[1][:3, :4]), 2.0, 6.0, device=dev)
sel = torch.from_numpy(np.random.default_rng(0).choice(H * W, n_rays, replace=False)).to(dev)
br = torch.stack([rays_all[sel][:, 0:3], rays_all[sel][:, 3:6]], 0).contiguous()
tgt = torch.rand(n_rays, 3, device=dev)
params = list(kw["network_fn"].parameters()) + list(kw["network_fine"].parameters())
opt = nb.Adam(params, lr=5e-4, betas=(0.9, 0.999))
pex = nd.PeerAdam([kw["network_fn"], kw["network_fine"]], opt, dev)
stepper = ntrain.GraphedTrainStep(n_rays, H, W, K, 32768, dict(kw, perturb=1.0), opt, 5e-4, 250, near=2.0, far=6.0, device=dev, exchange=pex)
for i in range(5):
    stepper(br, tgt, i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(iters):
    stepper(br, tgt, 5 + i)
e1.record()
torch.cuda.synchronize()
print(f"graphed step with PeerAdam, {n_rays} rays, world 1: {e0.elapsed_time(e1) / iters:.4f} ms per step")
