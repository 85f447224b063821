"""8-NN precompute of one 800x800 view against P=3 rendered base views: brute force vs grid (thread / warp kernels)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerfail_b200 as nb
from nerfail_b200 import ops, pipeline
from oracle import synth
from bench import LegoArgs, H, W

dev = torch.device("cuda:0")
_, kw, *_ = nb.create_nerf(LegoArgs(), device=dev)
kw["network_fn"].load_state_dict(synth.make_non_degenerate(synth.random_state_dict(0), 0))
kw["network_fine"].load_state_dict(synth.make_non_degenerate(synth.random_state_dict(1), 1))
K, _ = synth.intrinsics(H, W)
poses = synth.camera_ring(8)
kwr = dict(kw, near=2.0, far=6.0)
base = torch.stack([pipeline.render_points(H, W, K, torch.tensor(poses[i][:3, :4]), 1024, **kwr) for i in (0, 3, 5)], 0)
qpts = pipeline.render_points(H, W, K, torch.tensor(poses[1][:3, :4]), 1024, **kwr).reshape(-1, 3)
cand = base.reshape(-1, 3).contiguous()
ev = lambda: torch.cuda.Event(enable_timing=True)

def timed(fn, n):
    r = fn(); torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    for _ in range(n):
        r = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n, r

ms_build, grid = timed(lambda: ops.KnnGrid(cand), 3)
stats = torch.zeros(1, dtype=torch.int64, device=dev)
ms_grid, (d_g, i_g) = timed(lambda: grid.query(qpts, stats), 5)
evals = float(stats.item()) / 6
print(f"grid ({os.environ.get('NERFAIL_B200_KNN_GRID', 'warp')}): build {ms_build:.2f} ms, query {ms_grid:.2f} ms, {evals / qpts.shape[0]:.0f} evals/query, "
      f"pruning {qpts.shape[0] * cand.shape[0] / evals:.0f}x, h = {grid.h:.4f}")
if os.environ.get("BRUTE", "1") == "1":
    ms_brute, (d_b, i_b) = timed(lambda: ops.knn8(qpts, cand), 1)
    print(f"brute force: {ms_brute:.1f} ms; identical: {bool(torch.equal(i_g, i_b) and torch.equal(d_g, d_b))}; speedup {ms_brute / ms_grid:.1f}x")
