"""Host-side cost of one bf16 retraining step: small batch (GPU not the bottleneck), wall clock per step and cProfile."""
import cProfile, os, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["NERFAIL_B200_TRAIN"] = "bf16"
import nerfail_b200 as nb
from nerfail_b200 import ops
from oracle import synth
from bench import LegoArgs

dev = torch.device("cuda:0")
R = int(os.environ.get("R", 512))
_, kw, _, grad_vars, opt = nb.create_nerf(LegoArgs(), device=dev)
kw["network_fn"].load_state_dict(synth.make_non_degenerate(synth.random_state_dict(0), 0))
kw["network_fine"].load_state_dict(synth.make_non_degenerate(synth.random_state_dict(1), 1))
K, _ = synth.intrinsics(800, 800)
rays_all = ops.get_ray_batch(800, 800, K, torch.tensor(synth.camera_ring(8)[1][:3, :4]), 2.0, 6.0, device=dev)
sel = torch.from_numpy(np.random.default_rng(0).choice(640000, R, replace=False)).to(dev)
rays = rays_all[sel].contiguous()
batch_rays = torch.stack([rays[:, 0:3], rays[:, 3:6]], 0)
target = torch.rand(R, 3, device=dev)
kwt = dict(kw, near=2.0, far=6.0, perturb=1.0)

def step(i):
    return nb.train_step(batch_rays, target, 800, 800, K, 32768, kwt, opt, 5e-4, 250, i)

for i in range(5): step(i)
torch.cuda.synchronize()
n = 50
t0 = time.perf_counter()
for i in range(n): step(i)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"R={R}: host {1e3 * (t1 - t0) / n:.3f} ms/step to enqueue, {1e3 * (t2 - t0) / n:.3f} ms/step including the GPU tail")
pr = cProfile.Profile()
pr.enable()
for i in range(20): step(i)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
