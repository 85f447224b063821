"""The oracle (oracle/*.py) against golden vectors minted from the unmodified reference
(tests/golden/make_golden.py).  CPU only.  Because both sides are torch-CPU fp32 running the same sequence of
ATen ops, agreement is expected to be bit-exact; the tolerances below are 0 wherever that holds."""
import numpy as np
import torch

from conftest import golden
from oracle import gauss_oracle as go
from oracle import nerf_oracle as no
from oracle import synth

T = torch.from_numpy


def same(a, b, tol=0.0):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    if tol == 0.0:
        assert np.array_equal(a, b, equal_nan=True), f"max abs diff {np.nanmax(np.abs(a - b))}"
    else:
        np.testing.assert_allclose(a, b, rtol=tol, atol=tol, equal_nan=True)


def test_positional_encoding_and_mlp():
    g = golden("mlp.npz")
    x = T(g["x"])
    same(no.positional_encoding(x, 10).numpy(), g["enc10"])
    same(no.positional_encoding(x, 4).numpy(), g["enc4"])
    sd = synth.make_non_degenerate(synth.random_state_dict(0), 0)
    with torch.no_grad():
        out = no.nerf_mlp(sd, T(g["feats"]))
    same(out.numpy(), g["out"], tol=1e-6)


def test_composite_forward_and_backward():
    g = golden("composite.npz")
    raw, z, rd = T(g["raw"]), T(g["z"]), T(g["rays_d"])
    for white in (False, True):
        o = no.composite(raw, z, rd, white_bkgd=white)
        for name, t in zip(("rgb", "disp", "acc", "weights", "depth"), o):
            same(t.numpy(), g[f"{name}_w{int(white)}"])
    # the no-density rays reproduce the reference's 0/0 -> NaN disparity
    assert np.isnan(g["disp_w0"][:4]).all()
    rg = raw[8:].clone().requires_grad_(True)
    o = no.composite(rg, z[8:], rd[8:], white_bkgd=True)
    sum((a * T(g[f"cot{i}"])).sum() for i, a in enumerate(o)).backward()
    same(rg.grad.numpy(), g["g_raw"], tol=1e-6)


def test_sample_pdf():
    g = golden("sample_pdf.npz")
    bins, w = T(g["bins"]), T(g["weights"])
    same(no.inverse_cdf_samples(bins, w, 128).numpy(), g["det"])
    same(no.inverse_cdf_samples(bins, w, 128, T(g["u_rnd"])).numpy(), g["rnd"])


def test_rays_and_full_render():
    g = golden("render.npz")
    H, W = int(g["H"]), int(g["W"])
    rays = no.camera_rays(H, W, g["K"], T(g["c2w"]), 2.0, 6.0)
    same(rays[:, 0:3].reshape(H, W, 3).numpy(), g["rays_o"])
    same(rays[:, 3:6].reshape(H, W, 3).numpy(), g["rays_d"])
    sd_c = synth.make_non_degenerate(synth.random_state_dict(0), 0)
    sd_f = synth.make_non_degenerate(synth.random_state_dict(1), 1)
    with torch.no_grad():
        out = no.render_image(H, W, g["K"], T(g["c2w"]), sd_c, sd_f, chunk=64, retraw=True)
    for k_or, k_g in (("rgb_map", "rgb"), ("disp_map", "disp"), ("acc_map", "acc"), ("pts_max", "pts_max"), ("raw", "raw"),
                      ("rgb0", "rgb0"), ("disp0", "disp0"), ("acc0", "acc0"), ("z_std", "z_std")):
        same(out[k_or].numpy(), g[k_g], tol=2e-6)


def test_stochastic_render_matches_reference_pytest_hook():
    g = golden("render_stochastic.npz")
    rays = T(g["rays"])
    sd_c = synth.make_non_degenerate(synth.random_state_dict(0), 0)
    sd_f = synth.make_non_degenerate(synth.random_state_dict(1), 1)
    # the reference's pytest=True hook re-seeds numpy with 0 before each draw (run_nerf.py:374-377, :288-291,
    # run_nerf_helpers.py:215-223)
    np.random.seed(0); t_rand = torch.Tensor(np.random.rand(rays.shape[0], 64))
    np.random.seed(0); u = torch.Tensor(np.random.rand(rays.shape[0], 128))
    o, d, v = rays[:, 0:3], rays[:, 3:6], rays[:, 8:11]
    with torch.no_grad():
        z = no.coarse_depths(rays, 64, False, t_rand)
        raw = no.query_network(sd_c, o[:, None] + d[:, None] * z[..., None], v)
        np.random.seed(0); n0 = torch.Tensor(np.random.rand(*raw[..., 3].shape) * 1.0)
        rgb0, disp0, acc0, w0, _ = no.composite(raw, z, d, True, noise=n0)
        zf, z_new, z_std = no.hierarchical_depths(z, w0, 128, u)
        rawf = no.query_network(sd_f, o[:, None] + d[:, None] * zf[..., None], v)
        np.random.seed(0); n1 = torch.Tensor(np.random.rand(*rawf[..., 3].shape) * 1.0)
        rgb, disp, acc, w, _ = no.composite(rawf, zf, d, True, noise=n1)
    same(rgb0.numpy(), g["rgb0"], tol=2e-6)
    same(z_std.numpy(), g["z_std"], tol=2e-6)
    same(rgb.numpy(), g["rgb_map"], tol=2e-6)
    same(acc.numpy(), g["acc_map"], tol=2e-6)


def test_training_step_gradients():
    g = golden("train_step.npz")
    rays, target = T(g["rays"]), T(g["target"])
    sd_c = {k: v.clone().requires_grad_(True) for k, v in synth.make_non_degenerate(synth.random_state_dict(0), 0).items()}
    sd_f = {k: v.clone().requires_grad_(True) for k, v in synth.make_non_degenerate(synth.random_state_dict(1), 1).items()}
    np.random.seed(0); t_rand = torch.Tensor(np.random.rand(rays.shape[0], 64))      # the reference's pytest hook
    np.random.seed(0); u = torch.Tensor(np.random.rand(rays.shape[0], 128))
    out = no.render_ray_batch(rays, sd_c, sd_f, t_rand=t_rand, u=u)
    loss = torch.mean((out["rgb_map"] - target) ** 2) + torch.mean((out["rgb0"] - target) ** 2)
    loss.backward()
    same(loss.detach().numpy(), g["loss"], tol=1e-6)
    for tag, sd in (("c", sd_c), ("f", sd_f)):
        for name, p in sd.items():
            key = f"{tag}.{name}"
            if key in g:
                np.testing.assert_allclose(p.grad.numpy(), g[key], rtol=1e-4, atol=1e-7)
            np.testing.assert_allclose(np.linalg.norm(p.grad.numpy().astype(np.float64)), g[f"norm.{key}"], rtol=1e-4, atol=1e-9)


def test_training_trajectory_with_adam_and_lr_decay():
    """oracle Trainer (R:776-800) against three optimisation steps of the unmodified reference (make_golden_train.py):
    the losses agree to 1e-5 and every weight tensor ends where the reference's did (Adam is torch's in both)."""
    g = golden("train_traj.npz")
    tr = no.Trainer(synth.make_non_degenerate(synth.random_state_dict(0), 0), synth.make_non_degenerate(synth.random_state_dict(1), 1))
    for i, gs in enumerate(g["global_steps"]):
        rays = no.rays_from_batch(T(g["batch_rays"][i]), 2.0, 6.0)
        np.random.seed(0); t_rand = torch.Tensor(np.random.rand(rays.shape[0], 64))      # the reference's pytest hook
        np.random.seed(0); u = torch.Tensor(np.random.rand(rays.shape[0], 128))
        loss = tr.step(rays, T(g["targets"][i]), int(gs), t_rand, u)
        assert abs(loss - g["losses"][i]) <= 1e-5 * abs(g["losses"][i]) + 1e-7, (i, loss, g["losses"][i])
    assert abs(tr.opt.param_groups[0]["lr"] - 5e-4 * 0.1 ** (2 / 250000)) < 1e-12
    init = {"c": synth.make_non_degenerate(synth.random_state_dict(0), 0), "f": synth.make_non_degenerate(synth.random_state_dict(1), 1)}
    for tag, sd in (("c", tr.sd_c), ("f", tr.sd_f)):
        for name, p in sd.items():
            key = f"{tag}.{name}"
            moved = float((p.detach() - init[tag][name]).norm())
            assert moved > 0, key
            if key in g:        # the update itself (3 Adam steps of ~lr each) matches to 1 % of its own size
                assert float((p.detach() - T(g[key])).norm()) < 1e-2 * moved, (key, moved)
            np.testing.assert_allclose(np.linalg.norm(p.detach().numpy().astype(np.float64)), g[f"norm.{key}"], rtol=1e-6)


def test_knife_edge_of_deterministic_sampling():
    """Why the fine pass of a deterministic render (perturb = 0) cannot be held to 1e-3 on EVERY pixel by any second
    implementation: the experiment on the reference's own arithmetic (the oracle is bit-exact against it, see above).
    Perturb the reference's coarse weights by ONE ulp (random direction) and redo resampling + fine pass:
      * with u = linspace(0, 1, 128) the last sample sits at u == 1.0 == cdf[-1] up to rounding and empty bins sit at the
        `denom < 1e-5` switch of run_nerf_helpers.py:237-238, so z_std changes on > 10 % of the rays — several times (measured 10x at 40x40) the rate
        of the stochastic path (random u) under the same perturbation;
      * the rendered colour then moves by > 1e-4 of its scale on some rays: a 6e-8 relative input change amplified > 1000x,
        i.e. to the order of the 1e-3 parity bar itself, while staying under the 3e-3 the GPU tests allow for such rays.
    tests/test_gpu_render.py / test_gpu_fullsize.py therefore gate the fine pass at ">= 95 % of the pixels within 1e-3,
    every pixel within 3e-3" and the coarse pass (no sampling decision upstream) at 1e-3 everywhere."""
    H = W = 24
    K, _ = synth.intrinsics(H, W)
    rays = no.camera_rays(H, W, K, torch.tensor(synth.pose_spherical(30.0, -30.0, 4.0)[:3, :4]), 2.0, 6.0)
    sd_c = synth.make_non_degenerate(synth.random_state_dict(0), 0)
    sd_f = synth.make_non_degenerate(synth.random_state_dict(1), 1)
    o, d, v = rays[:, 0:3], rays[:, 3:6], rays[:, 8:11]
    gen = torch.Generator().manual_seed(0)
    with torch.no_grad():
        z = no.coarse_depths(rays, 64)
        raw = no.query_network(sd_c, o[:, None] + d[:, None] * z[..., None], v)
        _, _, _, w0, _ = no.composite(raw, z, d, True)

        def fine(zf):
            return no.composite(no.query_network(sd_f, o[:, None] + d[:, None] * zf[..., None], v), zf, d, True)[0]

        rate, worst = {}, {}
        for name, u in (("det", None), ("rand", torch.rand(rays.shape[0], 128, generator=gen))):
            zf, _, zstd = no.hierarchical_depths(z, w0, 128, u)
            rgb = fine(zf)
            changed, moved = [], []
            for _ in range(3):
                up = torch.rand(w0.shape, generator=gen) < 0.5
                w1 = torch.where(up, torch.nextafter(w0, torch.full_like(w0, 2.0)), torch.nextafter(w0, torch.full_like(w0, -1.0)))
                zf1, _, zstd1 = no.hierarchical_depths(z, w1, 128, u)
                changed.append(float(((zstd1 - zstd).abs() > 1e-6).float().mean()))
                moved.append(float((fine(zf1) - rgb).abs().max() / rgb.abs().max()))
            rate[name], worst[name] = float(np.mean(changed)), float(np.max(moved))
    assert rate["det"] > 0.10, rate
    assert rate["det"] > 4 * rate["rand"], rate
    assert 1e-4 < worst["det"] < 3e-3, worst


def test_gaussnet():
    g = golden("gauss.npz")
    s, di, ori = T(g["spatial_rgb"]), T(g["dist_idx"]), T(g["ori"])
    i_w = go.gaussian_weights(di, 0.02)
    same(i_w.numpy(), g["i_w"])
    for tag, eps in (("none", None), ("e32", 32), ("e2", 2)):
        sg = s.clone().requires_grad_(True)
        x, x_rgba, ext = go.gauss_forward(sg, i_w, ori, eps)
        same(x.detach().numpy(), g[f"x_{tag}"], tol=1e-6)
        same(x_rgba.detach().numpy(), g[f"xrgba_{tag}"], tol=1e-5)
        assert abs(ext[0] - float(g[f"epsmin_{tag}"])) < 1e-4 and abs(ext[1] - float(g[f"epsmax_{tag}"])) < 1e-4
        ((x * T(g[f"cx_{tag}"])).sum() + (x_rgba * T(g[f"cr_{tag}"])).sum()).backward()
        np.testing.assert_allclose(sg.grad.numpy(), g[f"grad_{tag}"], rtol=1e-5, atol=1e-5)


def test_knn_exact_vs_reference_procedure():
    """Bit-exact parity is defined against the direct-difference oracle; the reference's own cdist-based
    procedure is compared statistically (its default matmul mode is numerically noisy)."""
    g = golden("knn.npz")
    q, c = g["query"].reshape(-1, 3), g["cand"]
    d, i = go.knn8_exact(q, c)
    i_direct = g["i_ref_direct"].reshape(-1, 8)
    i_mm = g["i_ref_mm"].reshape(-1, 8)
    agree_direct = (i == i_direct).all(axis=1).mean()
    agree_mm_set = np.mean([set(a) == set(b) for a, b in zip(i, i_mm)])
    assert agree_direct >= 0.99, agree_direct          # direct-difference cdist orders identically (up to exact ties)
    assert agree_mm_set >= 0.80, agree_mm_set
    np.testing.assert_allclose(d, g["d_ref_direct"].reshape(-1, 8), rtol=2e-6, atol=1e-7)
    # brute-force fp64 check of the exactness claim
    d64 = np.linalg.norm(q[:, None, :].astype(np.float64) - c[None].astype(np.float64), axis=-1)
    i64 = np.argsort(d64, axis=1, kind="stable")[:, :8]
    assert (i == i64).all(axis=1).mean() >= 0.99
