"""bf16 tensor-core training against the fp32 exact path on the same problem: a student NeRF is fitted to views rendered
from a teacher NeRF (synthetic, no dataset), same initial weights, same ray batches; prints the loss curves and the PSNR of a
held-out view rendered from each student.  python scripts/train_convergence.py [steps] [rays per step]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerfail_b200 as nb
from oracle import synth
from bench import LegoArgs


def run(precision, steps, n_rand, dev, H=64, W=64, n_views=6, log=True):
    os.environ["NERFAIL_B200_TRAIN"] = precision
    K, focal = synth.intrinsics(H, W)
    poses = np.stack(synth.camera_ring(n_views + 1)).astype(np.float32)
    _, kw_t, *_ = nb.create_nerf(LegoArgs(), device=dev)
    kw_t["network_fn"].load_state_dict(synth.make_non_degenerate(synth.random_state_dict(2), 2, target_std=0.5))
    kw_t["network_fine"].load_state_dict(synth.make_non_degenerate(synth.random_state_dict(3), 3, target_std=0.5))
    kwr = dict(kw_t, near=2.0, far=6.0)
    with torch.no_grad():
        images = torch.stack([nb.render(H, W, K, chunk=4096, c2w=torch.tensor(p[:3, :4]), **kwr)[0] for p in poses], 0)
    kw_s, kw_test, _, _, opt = nb.create_nerf(LegoArgs(), device=dev)
    kw_s["network_fn"].load_state_dict(synth.make_non_degenerate(synth.random_state_dict(4), 4, target_std=0.5))
    kw_s["network_fine"].load_state_dict(synth.make_non_degenerate(synth.random_state_dict(5), 5, target_std=0.5))
    kws = dict(kw_s, near=2.0, far=6.0)
    rng = np.random.RandomState(0)
    torch.manual_seed(0)
    losses = []
    for i in range(steps):
        rays, tgt, _, _ = nb.sample_ray_batch(images, poses, list(range(n_views)), H, W, K, n_rand, i, 0, 0.5, rng=rng, device=dev)
        out = nb.train_step(rays, tgt, H, W, K, 32768, kws, opt, 5e-4, 250, i)
        if i % 10 == 0 or i == steps - 1:
            losses.append(float(out["loss"]))
    with torch.no_grad():
        kwe = dict(kw_test, near=2.0, far=6.0)
        kwe["network_fn"], kwe["network_fine"] = kw_s["network_fn"], kw_s["network_fine"]
        rgb = nb.render(H, W, K, chunk=4096, c2w=torch.tensor(poses[n_views][:3, :4]), **kwe)[0]
        psnr = float(-10.0 * torch.log10(((rgb - images[n_views]) ** 2).mean()))
    if log:
        print(f"{precision}: loss every 10 steps {[round(l, 5) for l in losses]}  held-out PSNR {psnr:.3f} dB")
    return losses, psnr


if __name__ == "__main__":
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    n_rand = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    dev = torch.device("cuda:0")
    l16, p16 = run("bf16", steps, n_rand, dev)
    l32, p32 = run("fp32", steps, n_rand, dev)
    print(f"held-out PSNR after {steps} steps of {n_rand} rays: bf16 {p16:.3f} dB, fp32 {p32:.3f} dB, difference {p16 - p32:+.3f} dB")
