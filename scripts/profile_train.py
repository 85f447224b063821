"""One bf16 retraining step (4096 rays, coarse + fine) for launch-list profiling."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["NERFAIL_B200_TRAIN"] = os.environ.get("NERFAIL_B200_TRAIN", "bf16")
import nerfail_b200 as nb
from nerfail_b200 import ops
from oracle import synth
from bench import LegoArgs

dev = torch.device("cuda:0")
_, kw, *_ = nb.create_nerf(LegoArgs(), device=dev)
kw["network_fn"].load_state_dict(synth.make_non_degenerate(synth.random_state_dict(0), 0))
kw["network_fine"].load_state_dict(synth.make_non_degenerate(synth.random_state_dict(1), 1))
K, _ = synth.intrinsics(800, 800)
rays_all = ops.get_ray_batch(800, 800, K, torch.tensor(synth.camera_ring(8)[1][:3, :4]), 2.0, 6.0, device=dev)
sel = torch.from_numpy(np.random.default_rng(0).choice(640000, 4096, replace=False)).to(dev)
rays = rays_all[sel].contiguous()
target = torch.rand(4096, 3, device=dev)
kwt = {k: v for k, v in kw.items() if k not in ("use_viewdirs", "ndc")}
kwt.update(perturb=1.0)
params = list(kw["network_fn"].parameters()) + list(kw["network_fine"].parameters())

def step():
    for p in params: p.grad = None
    with torch.enable_grad():
        ret = nb.render_rays(rays, retraw=True, **kwt)
        loss = nb.img2mse(ret["rgb_map"], target) + nb.img2mse(ret["rgb0"], target)
        loss.backward()
    return loss

for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for _ in range(n): step()
e1.record(); torch.cuda.synchronize()
print(f"train step: {e0.elapsed_time(e1) / n:.3f} ms (GPU events), {(time.perf_counter() - t0) / n * 1e3:.3f} ms wall")
