"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
The reference's hot-path modules import imageio / matplotlib / wandb at module scope without using them in
any arithmetic; empty stub modules are registered for whichever of those is missing (SURVEY.md §8c recipe).
Inputs are seeded through oracle/synth.py so the GPU tests can rebuild them bit-identically, but they are
ALSO stored in the .npz files, so the fixtures are self-contained.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("NERFAIL_REFERENCE", "/root/reference")
OUT = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)

from oracle import synth  # noqa: E402


def _stub_missing(names):
    for n in names:
        try:
            importlib.import_module(n)
        except Exception:
            m = types.ModuleType(n)
            sys.modules[n] = m
            if "." in n:
                setattr(sys.modules[n.split(".")[0]], n.split(".")[1], m)


def import_reference():
    _stub_missing(["imageio", "matplotlib", "matplotlib.pyplot", "wandb", "configargparse"])
    sys.path.insert(0, os.path.join(REF, "Create_spatial_point_set", "nerf_pytorch"))
    sys.path.insert(0, os.path.join(REF, "Create_spatial_point_set"))
    sys.path.insert(0, REF)
    import run_nerf_helpers as helpers          # upstream helpers
    import run_nerf                              # upstream render path
    import nerf_to_coord                         # NeRFail's copy with pts_max
    from model import GaussNet
    return helpers, run_nerf, nerf_to_coord, GaussNet


def np_(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def build_ref_nets(helpers, sd_c, sd_f):
    def mk(sd):
        net = helpers.NeRF(D=8, W=256, input_ch=63, output_ch=5, skips=[4], input_ch_views=27, use_viewdirs=True)
        net.load_state_dict(sd)
        return net
    return mk(sd_c), mk(sd_f)


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    helpers, run_nerf, n2c, GaussNet = import_reference()
    os.makedirs(OUT, exist_ok=True)
    g = torch.Generator().manual_seed(7)

    # ---- 1. positional encoding + MLP forward -------------------------------------------------
    embed10, ch10 = helpers.get_embedder(10, 0)
    embed4, ch4 = helpers.get_embedder(4, 0)
    x = (torch.rand(37, 3, generator=g) * 8 - 4)
    sd = synth.make_non_degenerate(synth.random_state_dict(0), 0)
    sd_fine = synth.make_non_degenerate(synth.random_state_dict(1), 1)
    net_c, net_f = build_ref_nets(helpers, sd, sd_fine)
    feats = torch.cat([embed10(x), embed4(torch.nn.functional.normalize(torch.randn(37, 3, generator=g), dim=-1))], -1)
    with torch.no_grad():
        mlp_out = net_c(feats)
    np.savez(os.path.join(OUT, "mlp.npz"), x=np_(x), enc10=np_(embed10(x)), enc4=np_(embed4(x)), feats=np_(feats),
             out=np_(mlp_out))

    # ---- 2. raw2outputs -----------------------------------------------------------------------
    R, S = 33, 64
    raw = torch.randn(R, S, 4, generator=g) * 2.0
    raw[:4, :, 3] = -1.0                       # rays with no density at all (acc = 0, disp = NaN path)
    raw[4:8, 10, 3] = 50.0                     # an opaque sample
    z = torch.sort(torch.rand(R, S, generator=g) * 4 + 2, dim=-1).values
    rd = torch.randn(R, 3, generator=g)
    outs = {}
    for white in (False, True):
        o = run_nerf.raw2outputs(raw, z, rd, 0, white, pytest=False)
        for name, t in zip(("rgb", "disp", "acc", "weights", "depth"), o):
            outs[f"{name}_w{int(white)}"] = np_(t)
    # gradients of a scalar functional of all five outputs w.r.t. raw (white background)
    raw_g = raw.clone().requires_grad_(True)
    o = run_nerf.raw2outputs(raw_g[8:], z[8:], rd[8:], 0, True)
    cot = [torch.randn(t.shape, generator=g) for t in o]
    sum((a * b).sum() for a, b in zip(o, cot)).backward()
    np.savez(os.path.join(OUT, "composite.npz"), raw=np_(raw), z=np_(z), rays_d=np_(rd), g_raw=np_(raw_g.grad[8:]),
             **{f"cot{i}": np_(c) for i, c in enumerate(cot)}, **outs)

    # ---- 3. sample_pdf ------------------------------------------------------------------------
    bins = torch.sort(torch.rand(R, 63, generator=g) * 4 + 2, dim=-1).values
    w = torch.rand(R, 62, generator=g) ** 4
    w[:3] = 0.0                                # flat pdf
    w[3:6, 5:] = 0.0                           # mass concentrated in a few bins
    det = helpers.sample_pdf(bins, w, 128, det=True)
    rnd = helpers.sample_pdf(bins, w, 128, det=False, pytest=True)
    np.random.seed(0)
    u_rnd = torch.Tensor(np.random.rand(R, 128))
    np.savez(os.path.join(OUT, "sample_pdf.npz"), bins=np_(bins), weights=np_(w), det=np_(det), rnd=np_(rnd), u_rnd=np_(u_rnd))

    # ---- 4. rays + full render_rays (coarse + fine, pts_max) -----------------------------------
    H = W = 12
    K, focal = synth.intrinsics(H, W)
    c2w = torch.tensor(synth.pose_spherical(30.0, -30.0, 4.0)[:3, :4])
    ro, rdir = helpers.get_rays(H, W, K, c2w)
    query = lambda inputs, viewdirs, fn: run_nerf.run_network(inputs, viewdirs, fn, embed_fn=embed10, embeddirs_fn=embed4,
                                                              netchunk=1 << 16)
    kw = dict(network_query_fn=query, perturb=0., N_importance=128, network_fine=net_f, N_samples=64, network_fn=net_c,
              white_bkgd=True, raw_noise_std=0.)
    with torch.no_grad():
        rgb, disp, acc, extras = run_nerf.render(H, W, K, chunk=64, c2w=c2w, ndc=False, near=2., far=6., use_viewdirs=True,
                                                 retraw=True, **kw)
        rgb2, disp2, acc2, pts_max, extras2 = n2c.render(H, W, K, chunk=64, c2w=c2w, ndc=False, near=2., far=6.,
                                                         use_viewdirs=True, **kw)
    assert torch.equal(rgb, rgb2)
    np.savez(os.path.join(OUT, "render.npz"), H=H, W=W, K=K, c2w=np_(c2w), rays_o=np_(ro), rays_d=np_(rdir), rgb=np_(rgb),
             disp=np_(disp), acc=np_(acc), pts_max=np_(pts_max), raw=np_(extras["raw"]), rgb0=np_(extras["rgb0"]),
             disp0=np_(extras["disp0"]), acc0=np_(extras["acc0"]), z_std=np_(extras["z_std"]))

    # stochastic path through the reference's pytest hook (perturb = 1, raw_noise_std = 1)
    rays = torch.cat([ro.reshape(-1, 3), rdir.reshape(-1, 3), 2 * torch.ones(H * W, 1), 6 * torch.ones(H * W, 1),
                      torch.nn.functional.normalize(rdir.reshape(-1, 3), dim=-1)], -1)[:40]
    kws = dict(kw); kws.update(perturb=1., raw_noise_std=1.)
    with torch.no_grad():
        st = run_nerf.render_rays(rays, retraw=True, pytest=True, **kws)
    np.savez(os.path.join(OUT, "render_stochastic.npz"), rays=np_(rays), **{k: np_(v) for k, v in st.items()})

    # ---- 5. training-mode forward + backward (config 5 in miniature) ---------------------------
    Rt = 24
    rays_t = torch.cat([ro.reshape(-1, 3), rdir.reshape(-1, 3), 2 * torch.ones(H * W, 1), 6 * torch.ones(H * W, 1),
                        torch.nn.functional.normalize(rdir.reshape(-1, 3), dim=-1)], -1)[60:60 + Rt]
    target = torch.rand(Rt, 3, generator=g)
    for n in (net_c, net_f):
        n.zero_grad()
    # training samples stochastically (configs/lego.txt: perturb = 1.0); the reference's pytest hook pins the draws.
    # Random u also keeps the test off the u == 1.0 knife edge of searchsorted that the deterministic path sits on.
    kwt = dict(kw); kwt.update(perturb=1.)
    ret = run_nerf.render_rays(rays_t, retraw=True, pytest=True, **kwt)
    loss = helpers.img2mse(ret["rgb_map"], target) + helpers.img2mse(ret["rgb0"], target)
    loss.backward()
    grads = {}
    for tag, n in (("c", net_c), ("f", net_f)):
        for name, p in n.named_parameters():
            grads[f"{tag}.{name}"] = np_(p.grad)
    keep = {k: v for k, v in grads.items() if any(s in k for s in (
        "pts_linears.0.", "pts_linears.5.weight", "pts_linears.7.bias", "alpha_linear", "rgb_linear", "views_linears.0.bias",
        "feature_linear.bias"))}
    norms = {f"norm.{k}": np.float64(np.linalg.norm(v.astype(np.float64))) for k, v in grads.items()}
    np.savez_compressed(os.path.join(OUT, "train_step.npz"), rays=np_(rays_t), target=np_(target), loss=np_(loss),
                        rgb=np_(ret["rgb_map"]), rgb0=np_(ret["rgb0"]), **keep, **norms)

    # ---- 6. GaussNet --------------------------------------------------------------------------
    P, Hg, Wg, B = 3, 20, 24, 2
    s, dist_idx, ori = synth.gauss_inputs(3, P, Hg, Wg, B, locality=True)
    cg = GaussNet.create_gauss_w("cpu", 0.02)
    i_w, dist_out = cg(dist_idx)

    class Identity8(torch.nn.Module):
        def forward(self, x):
            return x.mean(dim=(2, 3))
    out = {}
    for tag, eps in (("none", None), ("e32", 32), ("e2", 2)):
        net = GaussNet.gauss_net("cpu", 0.02, Identity8(), "my_model", epsilon=eps)
        s_g = s.clone().requires_grad_(True)
        x, x_rgba, cla, ori_f, ori_cla = net(s_g, i_w, ori)
        cx, cr = torch.randn(x.shape, generator=g), torch.randn(x_rgba.shape, generator=g)
        ((x * cx).sum() + (x_rgba * cr).sum()).backward()
        out.update({f"x_{tag}": np_(x), f"xrgba_{tag}": np_(x_rgba), f"cx_{tag}": np_(cx), f"cr_{tag}": np_(cr),
                    f"grad_{tag}": np_(s_g.grad), f"epsmin_{tag}": net.epsilon_3d_min, f"epsmax_{tag}": net.epsilon_3d_max})
    np.savez(os.path.join(OUT, "gauss.npz"), spatial_rgb=np_(s), dist_idx=np_(dist_idx), ori=np_(ori), i_w=np_(i_w), **out)

    # ---- 7. 8-NN: the reference procedure on a small cloud (statistics only) -------------------
    gq = torch.Generator().manual_seed(11)
    cand = torch.rand(3000, 3, generator=gq) * 2 - 1
    qry = (cand[torch.randint(0, 3000, (10 * 16,), generator=gq)] + 0.004 * torch.randn(160, 3, generator=gq)).reshape(10, 16, 3)
    from oracle import gauss_oracle
    d_ref, i_ref = gauss_oracle.knn8_reference_style(qry, cand, chunk=1200)
    d_dir, i_dir = gauss_oracle.knn8_reference_style(qry, cand, chunk=1200, compute_mode="donot_use_mm_for_euclid_dist")
    np.savez(os.path.join(OUT, "knn.npz"), cand=np_(cand), query=np_(qry), d_ref_mm=np_(d_ref), i_ref_mm=np_(i_ref),
             d_ref_direct=np_(d_dir), i_ref_direct=np_(i_dir))
    print("golden vectors written to", OUT)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f"  {f}: {os.path.getsize(os.path.join(OUT, f)) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
