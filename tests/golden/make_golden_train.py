"""Generates tests/golden/train_traj.npz: three optimisation steps of the UNMODIFIED reference on the CPU.

    python tests/golden/make_golden_train.py        (build container only: needs /root/reference)

The reference's training step is inline code of train() (run_nerf.py:776-800), so the lines are driven here with the
reference's own pieces: run_nerf.render (rays=batch_rays, retraw=True, pytest=True pins the stratified draws),
run_nerf_helpers.img2mse, torch.optim.Adam(lr=5e-4, betas=(0.9, 0.999)) as create_nerf builds it (:213) and the
learning-rate update of :796-800.  Stored: the inputs of every step, the loss of every step and the final weights.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg          # noqa: E402  (import_reference, build_ref_nets, synth)
from oracle import synth          # noqa: E402


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    helpers, run_nerf, _, _ = mg.import_reference()
    sd_c = synth.make_non_degenerate(synth.random_state_dict(0), 0)
    sd_f = synth.make_non_degenerate(synth.random_state_dict(1), 1)
    net_c, net_f = mg.build_ref_nets(helpers, sd_c, sd_f)
    embed10, _ = helpers.get_embedder(10, 0)
    embed4, _ = helpers.get_embedder(4, 0)
    query = lambda inputs, viewdirs, network_fn: run_nerf.run_network(inputs, viewdirs, network_fn, embed_fn=embed10,
                                                                      embeddirs_fn=embed4, netchunk=1 << 16)
    kw = dict(network_query_fn=query, perturb=1., N_importance=128, network_fine=net_f, N_samples=64, network_fn=net_c,
              white_bkgd=True, raw_noise_std=0., ndc=False, near=2., far=6., use_viewdirs=True)
    H = W = 16
    K, _ = synth.intrinsics(H, W)
    lrate, lrate_decay = 5e-4, 250
    opt = torch.optim.Adam(list(net_c.parameters()) + list(net_f.parameters()), lr=lrate, betas=(0.9, 0.999))
    g = torch.Generator().manual_seed(21)
    poses = synth.camera_ring(4)
    steps, n_rand = 3, 24
    batches, targets, losses = [], [], []
    for i in range(steps):
        ro, rd = helpers.get_rays(H, W, K, torch.Tensor(poses[i][:3, :4]))
        sel = torch.randperm(H * W, generator=g)[:n_rand]
        batch_rays = torch.stack([ro.reshape(-1, 3)[sel], rd.reshape(-1, 3)[sel]], 0)
        target_s = torch.rand(n_rand, 3, generator=g)
        rgb, disp, acc, extras = run_nerf.render(H, W, K, chunk=1024, rays=batch_rays, retraw=True, pytest=True, **kw)   # :776
        opt.zero_grad()
        loss = helpers.img2mse(rgb, target_s) + helpers.img2mse(extras['rgb0'], target_s)                              # :781-789
        loss.backward()
        opt.step()
        new_lrate = lrate * (0.1 ** (i / (lrate_decay * 1000)))                # :796-798; global_step == i here (starts at 0 :657, += 1 at :888)
        for pg in opt.param_groups:
            pg['lr'] = new_lrate
        batches.append(mg.np_(batch_rays)); targets.append(mg.np_(target_s)); losses.append(float(loss))
    out = {"batch_rays": np.stack(batches), "targets": np.stack(targets), "losses": np.asarray(losses, np.float64),
           "global_steps": np.arange(steps)}
    for tag, n in (("c", net_c), ("f", net_f)):
        for name, p in n.named_parameters():
            if any(s in name for s in ("pts_linears.0.", "pts_linears.5.weight", "alpha_linear", "rgb_linear", "views_linears.0.bias")):
                out[f"{tag}.{name}"] = mg.np_(p)
            out[f"norm.{tag}.{name}"] = np.float64(np.linalg.norm(mg.np_(p).astype(np.float64)))
    np.savez_compressed(os.path.join(mg.OUT, "train_traj.npz"), **out)
    print("losses", losses)


if __name__ == "__main__":
    main()
