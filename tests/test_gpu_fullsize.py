"""BASELINE.json's configurations at their stated sizes, CUDA path (through the C ABI) against the CPU oracle.

  config 1  one 100x100 view, lego config, coarse + fine           -> whole image against oracle.render_image
  config 3  GaussNet gather / scatter at 800x800, P = 3, 2 views   -> against gauss_oracle + torch autograd
  config 5  4096-ray retraining batch, forward + backward           -> loss, images and all 48 parameter gradients against
                                                                       oracle.render_ray_batch + torch autograd on the CPU
  and the bf16 (tensor-core) training path trained for N optimisation steps against oracle.Trainer (the reference's loop,
  pinned by tests/golden/train_traj.npz), several seeds, held-out view within north_star's 0.05 dB.
Tolerances (north_star): fp32 path 1e-3 relative; bf16 path within 0.05 dB PSNR of the fp32 reference.
The oracle runs on the host cores of the GPU box: the whole file takes a few minutes.
"""
import os

import numpy as np
import pytest
import torch

from oracle import gauss_oracle as go
from oracle import nerf_oracle as no
from oracle import synth
from test_gpu_render import Args, make_kwargs, psnr

pytestmark = pytest.mark.gpu
NOISE_30DB = 10 ** (-30 / 10)          # variance of the modelled photo noise: a trained NeRF's ~30 dB test PSNR


def expected_psnr_loss(img, ref, s2=NOISE_30DB):
    """PSNR lost against a photo modelled as `ref` + independent noise of variance s2: E|r - p|^2 = |r - ref|^2 + s2."""
    mse = float(np.mean((np.asarray(img, np.float64) - np.asarray(ref, np.float64)) ** 2))
    return 10 * np.log10(1 + mse / s2)


# ------------------------------------------------------------------------------------------------------------------
# config 1
# ------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def config1():
    torch.set_num_threads(os.cpu_count() or 1)
    H = W = 100
    K, _ = synth.intrinsics(H, W)
    c2w = torch.tensor(synth.pose_spherical(30.0, -30.0, 4.0)[:3, :4])
    sd_c = synth.make_non_degenerate(synth.random_state_dict(0), 0)
    sd_f = synth.make_non_degenerate(synth.random_state_dict(1), 1)
    with torch.no_grad():
        ref = no.render_image(H, W, K, c2w, sd_c, sd_f, chunk=1024)
    return H, W, K, c2w, {k: v.numpy() for k, v in ref.items()}


def test_config1_100x100_view_fp32_path(cuda, config1, monkeypatch):
    """fp32 accumulate path, whole 100x100 image, chunk 1024 as BASELINE configs[0] states: the coarse pass (no sampling
    decision upstream) within 1e-3 relative everywhere; the fine pass within 1e-3 on >= 95 % of the pixels, within 3e-3 on
    >= 99.9 % and within 1e-2 everywhere — the deterministic sampler's u == 1.0 == cdf[-1] knife edge, which the reference
    itself shows under a 1-ulp perturbation of its own coarse weights (1e-3 of scale from a 6e-8 input change,
    tests/test_oracle_golden.py::test_knife_edge_of_deterministic_sampling); over 10 000 rays the tail of that amplification
    reaches a few 1e-3 (measured on B200: one disparity pixel of a nearly empty ray at 3.05e-3, everything else < 3e-3)."""
    import nerfail_b200 as nb
    monkeypatch.setenv("NERFAIL_B200_MLP", "fp32")
    H, W, K, c2w, ref = config1
    _, kw = make_kwargs(cuda)
    with torch.no_grad():
        rgb, disp, acc, pts_max, extras = nb.render(H, W, K, chunk=1024, c2w=c2w, near=2., far=6., with_pts_max=True, **kw)
    for got, key in ((extras["rgb0"], "rgb0"), (extras["disp0"], "disp0"), (extras["acc0"], "acc0")):
        err = np.abs(got.cpu().numpy() - ref[key]).max()
        assert err <= 1e-3 * np.abs(ref[key]).max(), (key, err)
    for got, key in ((rgb, "rgb_map"), (disp, "disp_map"), (acc, "acc_map"), (extras["z_std"], "z_std")):
        err = np.abs(got.cpu().numpy() - ref[key])
        if err.ndim == 3:
            err = err.max(-1)
        scale = np.abs(ref[key]).max()
        frac, frac3 = float((err <= 1e-3 * scale).mean()), float((err <= 3e-3 * scale).mean())
        print(f"config 1 fp32 {key}: within 1e-3 {frac:.4f}, within 3e-3 {frac3:.5f}, max {float(err.max()) / scale:.2e} of scale")
        assert frac >= 0.95, (key, frac)
        assert frac3 >= 0.999, (key, frac3)
        assert err.max() <= 1e-2 * scale, (key, float(err.max()))
    same = float((np.abs(pts_max.cpu().numpy() - ref["pts_max"]).max(-1) < 1e-5).mean())
    assert same > 0.97, f"pts_max agrees on {same:.3f} of the pixels"


def test_config1_100x100_view_bf16_path(cuda, config1):
    """Default (bf16 tcgen05) path on the same image: PSNR loss <= 0.05 dB against the fp32 reference."""
    import nerfail_b200 as nb
    H, W, K, c2w, ref = config1
    _, kw = make_kwargs(cuda)
    with torch.no_grad():
        rgb, disp, acc, extras = nb.render(H, W, K, chunk=1024, c2w=c2w, near=2., far=6., **kw)
    kw["network_fn"].fused().status(); kw["network_fine"].fused().status()
    d = expected_psnr_loss(rgb.cpu().numpy(), ref["rgb_map"])
    d0 = expected_psnr_loss(extras["rgb0"].cpu().numpy(), ref["rgb0"])
    direct = psnr(rgb.cpu().numpy(), ref["rgb_map"])
    print(f"config 1 bf16: direct PSNR {direct:.2f} dB, expected loss at 30 dB: fine {d:.4f} dB, coarse {d0:.4f} dB")
    assert d < 0.05 and d0 < 0.05, (d, d0)
    assert direct > 40.0
    assert float(np.abs(acc.cpu().numpy() - ref["acc_map"]).mean()) < 1e-2


# ------------------------------------------------------------------------------------------------------------------
# config 3
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("eps", [32, None])
def test_config3_800x800_P3_two_views(cuda, eps):
    """GaussNet.py:53-119 at the attack's size: 3 x 800 x 800 x 4 perturbation table, two 800 x 800 views, Gaussian weights
    from distances (create_gauss_w), gather / composite forward and the scatter backward for upstream gradients on both
    outputs.  x bit-close (1e-5), x_rgba 1e-3, gradient w.r.t. the table 1e-3 of its scale."""
    import nerfail_b200 as nb
    torch.set_num_threads(os.cpu_count() or 1)
    P, H, W, B = 3, 800, 800, 2
    s, di, ori = synth.gauss_inputs(11, P, H, W, B, locality=True)
    g = torch.Generator().manual_seed(3)
    cx, cr = torch.randn(B, H, W, 4, generator=g), torch.randn(B, H, W, 4, generator=g)
    # oracle
    so = s.clone().requires_grad_(True)
    iw_o = go.gaussian_weights(di, 0.02)
    xo, xro, ext = go.gauss_forward(so, iw_o, ori, eps)
    ((xo * cx).sum() + (xro * cr).sum()).backward()
    # CUDA
    i_w, _ = nb.create_gauss_w(cuda, 0.02)(di.to(cuda))
    np.testing.assert_allclose(i_w.cpu().numpy(), iw_o.numpy(), rtol=2e-6, atol=1e-7)
    net = nb.gauss_net(cuda, 0.02, None, "my_model", epsilon=eps)
    sg = s.to(cuda).requires_grad_(True)
    x, x_rgba = net.perturbed(sg, i_w, ori.to(cuda))
    ((x * cx.to(cuda)).sum() + (x_rgba * cr.to(cuda)).sum()).backward()
    assert float((x.detach().cpu() - xo.detach()).abs().max()) <= 1e-5 * float(xo.detach().abs().max())
    assert float((x_rgba.detach().cpu() - xro.detach()).abs().max()) <= 1e-3
    gerr = float((sg.grad.cpu() - so.grad).abs().max())
    assert gerr <= 1e-3 * float(so.grad.abs().max()), gerr
    rel = float((sg.grad.cpu() - so.grad).norm() / so.grad.norm())
    assert rel < 1e-5, rel
    assert abs(net.epsilon_3d_min - ext[0]) <= 1e-4 * max(1.0, abs(ext[0])) and abs(net.epsilon_3d_max - ext[1]) <= 1e-4 * max(1.0, abs(ext[1]))


# ------------------------------------------------------------------------------------------------------------------
# config 5
# ------------------------------------------------------------------------------------------------------------------
N_RAND = 4096


@pytest.fixture(scope="module")
def config5():
    """The 4096-ray batch of BASELINE configs[4] (np.random.choice of the 800x800 pixels without replacement,
    run_nerf.py:768; targets U[0,1)^3) and the oracle's loss / images / parameter gradients for it, accumulated over
    512-ray pieces on the CPU (the loss is a mean over rays, so the pieces' gradients add)."""
    torch.set_num_threads(os.cpu_count() or 1)
    H = W = 800
    K, _ = synth.intrinsics(H, W)
    c2w = torch.tensor(synth.camera_ring(8)[1][:3, :4])
    rays_all = no.camera_rays(H, W, K, c2w, 2.0, 6.0)
    sel = np.random.default_rng(0).choice(H * W, N_RAND, replace=False)
    rays = rays_all[torch.from_numpy(sel)].contiguous()
    target = torch.rand(N_RAND, 3, generator=torch.Generator().manual_seed(5))
    np.random.seed(0); t_rand = torch.Tensor(np.random.rand(N_RAND, 64))          # the reference's pytest hook (run_nerf.py:374-377)
    np.random.seed(0); u = torch.Tensor(np.random.rand(N_RAND, 128))              # run_nerf_helpers.py:215-223
    sd_c = {k: v.clone().requires_grad_(True) for k, v in synth.make_non_degenerate(synth.random_state_dict(0), 0).items()}
    sd_f = {k: v.clone().requires_grad_(True) for k, v in synth.make_non_degenerate(synth.random_state_dict(1), 1).items()}
    loss_total, rgb, rgb0 = 0.0, [], []
    for i in range(0, N_RAND, 512):
        sl = slice(i, i + 512)
        out = no.render_ray_batch(rays[sl], sd_c, sd_f, t_rand=t_rand[sl], u=u[sl])
        part = (torch.sum((out["rgb_map"] - target[sl]) ** 2) + torch.sum((out["rgb0"] - target[sl]) ** 2)) / (3 * N_RAND)
        part.backward()
        loss_total += float(part)
        rgb.append(out["rgb_map"].detach()); rgb0.append(out["rgb0"].detach())
    grads = {"c": {k: v.grad.numpy().astype(np.float64) for k, v in sd_c.items()},
             "f": {k: v.grad.numpy().astype(np.float64) for k, v in sd_f.items()}}
    return rays, target, loss_total, torch.cat(rgb).numpy(), torch.cat(rgb0).numpy(), grads


def _gpu_step(cuda, rays, target):
    import nerfail_b200 as nb
    nets = []
    for seed in (0, 1):
        n = nb.NeRF(D=8, W=256, input_ch=63, output_ch=5, skips=[4], input_ch_views=27, use_viewdirs=True).to(cuda)
        n.load_state_dict(synth.make_non_degenerate(synth.random_state_dict(seed), seed))
        nets.append(n)
    e10, _ = nb.get_embedder(10)
    e4, _ = nb.get_embedder(4)
    query = nb.NetworkQuery(e10, e4, 1 << 16)
    tgt = target.to(cuda)
    ret = nb.render_rays(rays.to(cuda), nets[0], query, 64, retraw=True, perturb=1., N_importance=128, network_fine=nets[1],
                         white_bkgd=True, raw_noise_std=0., pytest=True)
    loss = nb.img2mse(ret["rgb_map"], tgt) + nb.img2mse(ret["rgb0"], tgt)
    loss.backward()
    return nets, ret, float(loss)


def _grad_errors(nets, grads):
    """per tensor: (relative L2 error, cosine) against the oracle gradient; per network: the same over all tensors."""
    per, agg = {}, {}
    for tag, n in (("c", nets[0]), ("f", nets[1])):
        num = den = dot = nn = 0.0
        for name, p in n.named_parameters():
            a, b = p.grad.cpu().numpy().astype(np.float64), grads[tag][name]
            e2, b2, a2, ab = float(((a - b) ** 2).sum()), float((b ** 2).sum()), float((a ** 2).sum()), float((a * b).sum())
            per[f"{tag}.{name}"] = ((e2 / (b2 + 1e-300)) ** 0.5, ab / ((a2 * b2) ** 0.5 + 1e-300))
            num += e2; den += b2; dot += ab; nn += a2
        agg[tag] = ((num / den) ** 0.5, dot / (nn * den) ** 0.5)
    return per, agg


def test_config5_4096_ray_step_fp32_path(cuda, config5, monkeypatch):
    """Exact-parity training path on the full 4096-ray batch: loss 1e-4, images 1e-3, every parameter gradient within 1e-3
    norm-wise and 3e-3 in relative L2 error of the oracle's autograd (relu masks within rounding of zero flip between
    two fp32 evaluations, see tests/test_gpu_mlp.py::test_training_step_gradients_vs_reference_golden)."""
    monkeypatch.setenv("NERFAIL_B200_TRAIN", "fp32")
    rays, target, loss_ref, rgb_ref, rgb0_ref, grads = config5
    nets, ret, loss = _gpu_step(cuda, rays, target)
    assert abs(loss - loss_ref) <= 1e-4 * loss_ref, (loss, loss_ref)
    np.testing.assert_allclose(ret["rgb_map"].detach().cpu().numpy(), rgb_ref, rtol=1e-3, atol=2e-4)
    np.testing.assert_allclose(ret["rgb0"].detach().cpu().numpy(), rgb0_ref, rtol=1e-3, atol=2e-4)
    per, agg = _grad_errors(nets, grads)
    for key, (rel, cos) in per.items():
        tag, name = key.split(".", 1)
        gn, rn = float(np.linalg.norm(dict(nets[0 if tag == "c" else 1].named_parameters())[name].grad.cpu().numpy().astype(np.float64))), float(np.linalg.norm(grads[tag][name]))
        assert abs(gn - rn) <= 1e-3 * rn + 1e-12, (key, gn, rn)
        assert rel <= 3e-3, (key, rel)
    for tag in ("c", "f"):
        assert agg[tag][0] <= 1.5e-3, (tag, agg[tag])


def test_config5_4096_ray_step_bf16_path(cuda, config5, monkeypatch):
    """The default (bf16 tensor-core) training kernels on the same batch against the fp32 ORACLE (not an emulation):
    the rendered batch within north_star's 0.05 dB budget (measured 0.018 / 0.004 dB), the loss within 1e-3 (measured 1e-4),
    and the parameter gradients — relative L2 error <= 1 % and cosine >= 0.9999 per network (measured 0.36 % / 0.999995),
    <= 8 % and >= 0.998 for every single tensor (measured worst: pts_linears.0.weight of the fine network, 4.6 % / 0.99898;
    mixed precision: bf16 operands, fp32 accumulation).  The values are printed."""
    monkeypatch.setenv("NERFAIL_B200_TRAIN", "bf16")
    rays, target, loss_ref, rgb_ref, rgb0_ref, grads = config5
    nets, ret, loss = _gpu_step(cuda, rays, target)
    for n in nets:
        n.fused().status()
    d = expected_psnr_loss(ret["rgb_map"].detach().cpu().numpy(), rgb_ref)
    d0 = expected_psnr_loss(ret["rgb0"].detach().cpu().numpy(), rgb0_ref)
    per, agg = _grad_errors(nets, grads)
    worst = sorted(per.items(), key=lambda kv: -kv[1][0])[:5]
    print(f"config 5 bf16: loss {loss:.6f} vs oracle {loss_ref:.6f}; PSNR budget fine {d:.4f} coarse {d0:.4f} dB; "
          f"gradient rel-L2 / cosine per network: coarse {agg['c'][0]:.4f} / {agg['c'][1]:.6f}, fine {agg['f'][0]:.4f} / {agg['f'][1]:.6f}; "
          f"worst tensors {[(k, round(v[0], 4), round(v[1], 5)) for k, v in worst]}")
    assert d < 0.05 and d0 < 0.05, (d, d0)
    assert abs(loss - loss_ref) <= 1e-3 * loss_ref, (loss, loss_ref)
    for tag in ("c", "f"):
        assert agg[tag][1] >= 0.9999 and agg[tag][0] <= 0.01, (tag, agg[tag])
    for key, (rel, cos) in per.items():
        assert cos >= 0.998 and rel <= 0.08, (key, rel, cos)


# ------------------------------------------------------------------------------------------------------------------
# N optimisation steps: bf16 tensor-core training against the oracle's (= the reference's) training loop
# ------------------------------------------------------------------------------------------------------------------
TRAIN_SEEDS = tuple(int(s) for s in os.environ.get("NERFAIL_TEST_TRAIN_SEEDS", "2,3,4,5,6,7").split(","))
TRAIN_STEPS = int(os.environ.get("NERFAIL_TEST_TRAIN_STEPS", "60"))
TRAIN_RAYS = int(os.environ.get("NERFAIL_TEST_TRAIN_RAYS", "256"))
RESULTS = []
TRAIN_SIGMA_STD = float(os.environ.get("NERFAIL_TEST_TRAIN_SIGMA_STD", "2.0"))      # W-B normalisation of teacher and student


def _train_pair(cuda, seed, steps, n_rand, Hs=32, n_views=4):
    """One fitting problem solved three times from the same state: by train_step on the GPU with the bf16 tensor-core
    kernels (the default), by train_step with the fp32 layer kernels (the 1e-3 parity path: it shows how far two fp32
    implementations drift apart on this problem), and by oracle.Trainer on the CPU (the reference's loop).  Same ray
    batches, same stratified draws (the reference's pytest hook everywhere).  All three weight sets are rendered on the
    held-out view by the SAME fp32 renderer (the oracle); the bf16-trained ones also by the bf16 kernels (the product)."""
    import nerfail_b200 as nb
    K, _ = synth.intrinsics(Hs, Hs)
    poses = np.stack(synth.camera_ring(n_views + 1)).astype(np.float32)
    t_c = synth.make_non_degenerate(synth.random_state_dict(100 + 2 * seed), 100 + 2 * seed, target_std=TRAIN_SIGMA_STD)
    t_f = synth.make_non_degenerate(synth.random_state_dict(101 + 2 * seed), 101 + 2 * seed, target_std=TRAIN_SIGMA_STD)
    with torch.no_grad():                                      # the "photographs": views of a teacher NeRF (fp32 oracle)
        images = torch.stack([no.render_image(Hs, Hs, K, torch.tensor(p[:3, :4]), t_c, t_f, chunk=1024)["rgb_map"] for p in poses], 0)
    s_c = synth.make_non_degenerate(synth.random_state_dict(200 + 2 * seed), 200 + 2 * seed, target_std=TRAIN_SIGMA_STD)
    s_f = synth.make_non_degenerate(synth.random_state_dict(201 + 2 * seed), 201 + 2 * seed, target_std=TRAIN_SIGMA_STD)
    rng = np.random.RandomState(seed)
    batches = [nb.sample_ray_batch(images, poses, list(range(n_views)), Hs, Hs, K, n_rand, i, 0, 0.5, rng=rng, device=cuda)[:2]
               for i in range(steps)]
    np.random.seed(0); t_rand = torch.Tensor(np.random.rand(n_rand, 64))
    np.random.seed(0); u = torch.Tensor(np.random.rand(n_rand, 128))
    held = torch.tensor(poses[n_views][:3, :4])
    photo = images[n_views].numpy()

    oracle = no.Trainer(s_c, s_f, 5e-4, 250)
    l_ref = [oracle.step(no.rays_from_batch(rays.cpu(), 2.0, 6.0), tgt.cpu(), i, t_rand, u) for i, (rays, tgt) in enumerate(batches)]
    with torch.no_grad():
        rgb_ref = no.render_image(Hs, Hs, K, held, *oracle.state_dicts(), chunk=1024)["rgb_map"].numpy()
    out = {"psnr_ref": psnr(rgb_ref, photo), "l_ref": l_ref, "degenerate": float(np.std(rgb_ref)) < 1e-3}
    prev = os.environ.get("NERFAIL_B200_TRAIN")
    try:
        for mode in ("bf16", "fp32"):
            os.environ["NERFAIL_B200_TRAIN"] = mode
            kw_train, kw_test, _, _, opt = nb.create_nerf(Args(), device=cuda)
            kw_train["network_fn"].load_state_dict(s_c); kw_train["network_fine"].load_state_dict(s_f)
            kws = dict(kw_train, near=2.0, far=6.0, pytest=True)
            losses = [float(nb.train_step(rays, tgt, Hs, Hs, K, 32768, kws, opt, 5e-4, 250, i)["loss"]) for i, (rays, tgt) in enumerate(batches)]
            with torch.no_grad():
                g_c = {k: v.detach().cpu() for k, v in kw_train["network_fn"].state_dict().items()}
                g_f = {k: v.detach().cpu() for k, v in kw_train["network_fine"].state_dict().items()}
                rgb_w = no.render_image(Hs, Hs, K, held, g_c, g_f, chunk=1024)["rgb_map"].numpy()      # these weights, fp32 oracle render
                out[f"psnr_weights_{mode}"] = psnr(rgb_w, photo)
                out[f"l_{mode}"] = losses
                if mode == "bf16":
                    rgb_prod = nb.render(Hs, Hs, K, chunk=4096, c2w=held, **dict(kw_test, near=2.0, far=6.0))[0].cpu().numpy()
                    out.update(psnr_product=psnr(rgb_prod, photo), budget_product=expected_psnr_loss(rgb_prod, rgb_ref),
                               direct_product=psnr(rgb_prod, rgb_ref))
    finally:
        if prev is None:
            os.environ.pop("NERFAIL_B200_TRAIN", None)
        else:
            os.environ["NERFAIL_B200_TRAIN"] = prev
    return out


def test_bf16_training_matches_oracle_training_within_psnr_budget(cuda, monkeypatch):
    """north_star's bf16 bar on the TRAINING path, N optimisation steps from identical states, several seeds:
      (a) the weights the tensor-core kernels produce (forward / data-gradient / weight-gradient in bf16, fused Adam) are as
          good as the oracle's fp32-trained ones: both rendered by the SAME fp32 renderer (the oracle) on a held-out view,
          PSNR against the target within 0.05 dB for every seed and on average (the fp32 layer kernels are trained next to
          them and printed: they show the noise floor of two fp32 trajectories on the same problem);
      (b) the product end to end (bf16-trained weights rendered by the bf16 kernels) stays within the 0.05 dB budget of the
          fp32 reference end to end (oracle-trained, oracle-rendered), measured like the inference tests: expected PSNR loss
          against a photograph at a 30 dB level;
      (c) the loss curves agree within 5 % at every step."""
    torch.set_num_threads(os.cpu_count() or 1)
    d_w, d_32, d_p, problems = [], [], [], []
    for seed in TRAIN_SEEDS:
        r = _train_pair(cuda, seed, TRAIN_STEPS, TRAIN_RAYS)
        l_gpu, l_ref = r["l_bf16"], r["l_ref"]
        dw, d32, dp = r["psnr_weights_bf16"] - r["psnr_ref"], r["psnr_weights_fp32"] - r["psnr_ref"], r["psnr_product"] - r["psnr_ref"]
        print(f"seed {seed}: held-out PSNR oracle-trained {r['psnr_ref']:.4f} dB; same renderer: bf16-trained {dw:+.4f} dB, fp32-kernel-trained {d32:+.4f} dB; "
              f"product (bf16-trained, bf16-rendered) {dp:+.4f} dB, vs reference render direct {r['direct_product']:.2f} dB = budget at 30 dB "
              f"{r['budget_product']:.4f} dB; loss first/last bf16 {l_gpu[0]:.5f}/{l_gpu[-1]:.5f} oracle {l_ref[0]:.5f}/{l_ref[-1]:.5f}"
              + ("  [degenerate: constant image]" if r["degenerate"] else ""))
        if not (np.isfinite(l_ref).all() and np.mean(l_ref[-10:]) < 1.5 * np.mean(l_ref[:10])):      # both optimisers are stable
            problems.append((seed, "oracle loss unstable"))
        if not np.allclose(l_gpu, l_ref, rtol=5e-2):      # per-step losses on 256 rays; measured worst 2.5 % (at a loss of 0.007)
            problems.append((seed, "loss curves differ by", float(np.abs(np.array(l_gpu) / np.array(l_ref) - 1).max())))
        if abs(r["budget_product"]) > 0.05:
            problems.append((seed, "product render outside the 0.05 dB budget", r["budget_product"]))
        if r["degenerate"]:           # the optimisation emptied the scene (all three implementations agree bit for bit): uninformative
            continue
        d_w.append(dw); d_32.append(d32); d_p.append(dp)
        RESULTS.append((seed, dw, d32, dp, r["budget_product"]))
    print(f"mean over {len(d_w)} seeds: bf16-trained {float(np.mean(d_w)):+.4f} dB (rms {float(np.sqrt(np.mean(np.square(d_w)))):.4f}), "
          f"fp32-kernel-trained {float(np.mean(d_32)):+.4f} dB (rms {float(np.sqrt(np.mean(np.square(d_32)))):.4f}), product {float(np.mean(d_p)):+.4f} dB")
    assert not problems, problems
    assert len(d_w) >= 4, "too few non-degenerate fitting problems"
    assert abs(float(np.mean(d_w))) <= 0.05, d_w
    assert max(abs(d) for d in d_w) <= 0.05, d_w        # measured on B200: rms 0.011 dB, worst seed 0.025 dB (fp32 kernels: 0.001 dB)
