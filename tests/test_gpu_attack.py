"""GPU parity of the attack-loop pieces around GaussNet: the fused classifier-input Resize and
its adjoint, batched per-class gradients (DeepFool), the sharded 8-NN sweep, the asynchronous RGBA PNG sink, and the
kernel-side Philox draws of the stochastic render path.  Oracle: oracle/gauss_oracle.py + torchvision / torch autograd on the
CPU (the reference's own third-party ops for Resize and the classifier)."""
import os

import numpy as np
import pytest
import torch

from oracle import gauss_oracle as go
from oracle import nerf_oracle as no
from oracle import synth
from philox_ref import uniform as philox_uniform_ref
from test_gpu_render import make_kwargs

pytestmark = pytest.mark.gpu


def reference_classifier_input(x_rgba, size, antialias):
    """GaussNet.py:121-154 with torch / torchvision on the CPU."""
    from torchvision.transforms import Resize
    c = x_rgba.transpose(2, 3).transpose(1, 2)
    rgb = torch.where(c[:, 3:4].expand(-1, 3, -1, -1) > 0, c[:, :3], torch.ones_like(c[:, :3]) * 255)
    return rgb if size is None else Resize([size, size], antialias=antialias)(rgb)


class TinyClassifier(torch.nn.Module):
    """Stand-in for the victim classifier (a third-party model in the reference): conv / relu / pool / linear, 8 classes."""

    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(9)
        self.conv = torch.nn.Conv2d(3, 6, 5, stride=3)
        self.fc = torch.nn.Linear(6 * 4 * 4, 8)
        with torch.no_grad():
            for p in self.parameters():
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)

    def forward(self, x):
        h = torch.relu(self.conv(x / 255.0))
        h = torch.nn.functional.adaptive_avg_pool2d(h, 4)
        return self.fc(h.flatten(1))


@pytest.mark.parametrize("size,antialias,H,W", [(299, True, 800, 800), (299, False, 800, 800), (224, True, 800, 800),
                                                 (299, True, 120, 90), (32, False, 50, 64)])
def test_fused_resize_forward_adjoint_and_second_derivative(cuda, size, antialias, H, W):
    """nfb_rgba_to_chw_resized against where(alpha > 0, rgb, 255) + torchvision Resize (1e-3 of the 255 range; measured
    ~1e-5), its autograd backward against torch's, and the backward of the backward (the pair must stay differentiable
    twice for deepfool.py:76-77) against torch's double backward."""
    from nerfail_b200 import ops
    g = torch.Generator().manual_seed(H + size)
    B = 2
    img = torch.rand(B, H, W, 4, generator=g) * 255
    img[..., 3] = (torch.rand(B, H, W, generator=g) > 0.3).float() * 255          # holes: alpha == 0 -> white
    cot = torch.randn(B, 3, size, size, generator=g)
    # reference: values, gradient, and gradient of <grad, v> w.r.t. the cotangent
    a = img.clone().requires_grad_(True)
    ref = reference_classifier_input(a, size, antialias)
    c_ref = cot.clone().requires_grad_(True)
    (g_ref,) = torch.autograd.grad(ref, a, c_ref, create_graph=True)
    v = torch.randn(B, H, W, 4, generator=g)
    (gg_ref,) = torch.autograd.grad((g_ref * v).sum(), c_ref)
    # CUDA
    b = img.to(cuda).requires_grad_(True)
    out = ops.RgbaToChwResizedFn.apply(b, 255.0, size, antialias)
    c_gpu = cot.to(cuda).requires_grad_(True)
    (g_gpu,) = torch.autograd.grad(out, b, c_gpu, create_graph=True)
    (gg_gpu,) = torch.autograd.grad((g_gpu * v.to(cuda)).sum(), c_gpu)
    assert float((out.detach().cpu() - ref.detach()).abs().max()) <= 1e-3 * 255
    assert float((out.detach().cpu() - ref.detach()).abs().max()) <= 2e-4 * 255
    ge = float((g_gpu.detach().cpu() - g_ref.detach()).abs().max())
    assert ge <= 1e-3 * float(g_ref.detach().abs().max()), ge
    gge = float((gg_gpu.cpu() - gg_ref).abs().max())
    assert gge <= 1e-3 * float(gg_ref.abs().max()), gge
    # uint8 input (the original image): same kernel, no gradient
    u8 = img.round().clamp(0, 255).to(torch.uint8)
    ref_u8 = reference_classifier_input(u8.float(), size, antialias)
    got_u8 = ops.rgba_u8_to_chw_resized(u8.to(cuda), size, antialias)
    assert float((got_u8.cpu() - ref_u8).abs().max()) <= 2e-4 * 255


@pytest.mark.parametrize("model_name", ["inception_v3", "vit_b_16", "my_model"])
def test_gauss_net_forward_through_the_classifier(cuda, model_name):
    """gauss_net.forward with the Resize fused in (GaussNet.py:46-159) against the oracle's x / x_rgba followed by the
    reference's own torch ops (where, Resize, model) on the CPU: logits and the gradient of a cross-entropy loss w.r.t. the
    perturbation table (what attack_NeRFail_S.py:331-348 computes) within 1e-3."""
    import nerfail_b200 as nb
    P, H, W, B = 2, 96, 80, 3
    s, di, ori = synth.gauss_inputs(4, P, H, W, B, locality=True)
    size = None if model_name == "my_model" else (224 if model_name == "vit_b_16" else 299)
    model = TinyClassifier()
    label = torch.tensor([2, 2, 2])
    crit = torch.nn.CrossEntropyLoss()
    so = s.clone().requires_grad_(True)
    iw = go.gaussian_weights(di, 0.02)
    xo, xro, _ = go.gauss_forward(so, iw, ori, 32)
    cla_o = model(reference_classifier_input(xro, size, True))
    ori_cla_o = model(reference_classifier_input(ori.float(), size, True))
    crit(cla_o, label).backward()

    net = nb.gauss_net(cuda, 0.02, TinyClassifier().to(cuda), model_name, epsilon=32)
    net.resize_antialias = True
    sg = s.to(cuda).requires_grad_(True)
    i_w, _ = nb.create_gauss_w(cuda, 0.02)(di.to(cuda))
    x, x_rgba, cla, ori_f, ori_cla = net(sg, i_w, ori.to(cuda))
    crit(cla, label.to(cuda)).backward()
    assert float((cla.detach().cpu() - cla_o.detach()).abs().max()) <= 1e-3 * float(cla_o.detach().abs().max())
    assert float((ori_cla.detach().cpu() - ori_cla_o.detach()).abs().max()) <= 1e-3 * float(ori_cla_o.detach().abs().max())
    gerr = float((sg.grad.cpu() - so.grad).abs().max())
    assert gerr <= 1e-3 * float(so.grad.abs().max()), gerr


@pytest.mark.parametrize("model_name,B", [("inception_v3", 1), ("my_model", 2)])
def test_batched_class_gradients_equal_per_class_autograd(cuda, model_name, B):
    """gauss_net.class_gradients (one Resize-adjoint launch + one batched scatter launch for all classes) against the
    reference procedure of deepfool.py:72-86: one torch.autograd.grad per class through the whole forward, on the CPU."""
    import nerfail_b200 as nb
    P, H, W = 3, 72, 64
    s, di, ori = synth.gauss_inputs(6, P, H, W, B, locality=True)
    size = None if model_name == "my_model" else 299
    model = TinyClassifier()
    so = s.clone().requires_grad_(True)
    iw = go.gaussian_weights(di, 0.02)
    _, xro, _ = go.gauss_forward(so, iw, ori, 32)
    cla_o = model(reference_classifier_input(xro, size, True))
    classes = [0, 1, 3, 4, 5, 6, 7]
    ref = torch.stack([torch.autograd.grad(cla_o[:, k].sum(), so, retain_graph=True)[0] for k in classes], 0)

    net = nb.gauss_net(cuda, 0.02, TinyClassifier().to(cuda), model_name, epsilon=32)
    net.resize_antialias = True
    i_w, _ = nb.create_gauss_w(cuda, 0.02)(di.to(cuda))
    from nerfail_b200 import _lib
    l0 = _lib.launch_count()
    grads, cla = net.class_gradients(s.to(cuda), i_w, ori.to(cuda), classes)
    launches = _lib.launch_count() - l0
    assert grads.shape == (len(classes), P, H, W, 4)
    assert float((cla.cpu() - cla_o.detach()).abs().max()) <= 1e-3 * float(cla_o.detach().abs().max())
    err = float((grads.cpu() - ref).abs().max())
    assert err <= 1e-3 * float(ref.abs().max()), err
    # gather, conversion(+resize), its adjoint (one launch; NC launches of the plain adjoint for my_model), one scatter
    assert launches <= (4 if size is not None else 3 + len(classes)), launches


def test_knn_sweep_sharded_by_view(cuda, tmp_path):
    """pipeline.knn_sweep: two ranks' shards (i % 2) together equal the single-rank sweep, the files written by either are
    the reference's [2,H,W,8] float32 tensors (create_index_and_dist.py:148-163, tools/dist_to_weight.py:95-97), and every
    view's indices are the exact 8-NN of the oracle."""
    import nerfail_b200 as nb
    from nerfail_b200 import pipeline
    g = torch.Generator().manual_seed(2)
    P, H, W, V = 2, 20, 24, 5
    base = torch.rand(P, H, W, 3, generator=g)
    views = [base.reshape(-1, 3)[torch.randint(0, P * H * W, (H * W,), generator=g)].reshape(H, W, 3)
             + 0.01 * torch.randn(H, W, 3, generator=g) for _ in range(V)]
    paths = []
    for i, v in enumerate(views):                      # views come from disk like the reference's coords/NNN.npy
        p = tmp_path / f"{i:03d}.npy"
        pipeline.save_points_npy(str(p), v)
        paths.append(str(p))
    one = dict(nb.knn_sweep(paths, base.to(cuda), out_dir=str(tmp_path / "one"), rank=0, world_size=1))
    two = {}
    for r in range(2):
        two.update(dict(nb.knn_sweep(paths, base.to(cuda), out_dir=str(tmp_path / "two"), rank=r, world_size=2)))
    assert sorted(one) == sorted(two) == list(range(V))
    for i in range(V):
        assert torch.equal(one[i], two[i])
        for sub in ("index_and_dist", "index_and_weight"):
            a = torch.load(tmp_path / "one" / sub / f"{i}.pth")
            b = torch.load(tmp_path / "two" / sub / f"{i}.pth")
            assert a.dtype == torch.float32 and tuple(a.shape) == (2, H, W, 8) and torch.equal(a, b)
        assert torch.equal(torch.load(tmp_path / "one" / "index_and_weight" / f"{i}.pth"), one[i].cpu())
        d_ref, i_ref = go.knn8_exact(views[i].numpy().reshape(-1, 3), base.numpy().reshape(-1, 3))
        di = torch.load(tmp_path / "one" / "index_and_dist" / f"{i}.pth").numpy()
        assert np.array_equal(di[1].reshape(-1, 8).astype(np.int32), i_ref)
        assert np.array_equal(di[0].reshape(-1, 8), d_ref)


def test_attack_image_sink_writes_the_reference_files(cuda, tmp_path):
    """AttackImageSink against the reference's blocking writes (attack_NeRFail_S.py:394-403: cv2.imwrite of the float
    tensors): identical bytes for x_rgba, x (the perturbation image, negative and > 255 values saturate) and the original."""
    import cv2
    import nerfail_b200 as nb
    g = torch.Generator().manual_seed(8)
    B, H, W = 3, 40, 52
    x_rgba = (torch.rand(B, H, W, 4, generator=g) * 255)
    x_rgba[0, 0, :4, 0] = torch.tensor([0.5, 1.5, 2.5, 254.5])                    # ties round to even like cv2
    x = torch.randn(B, H, W, 4, generator=g) * 200
    ori = torch.randint(0, 256, (B, H, W, 4), generator=g).float()
    ref_dir, got_dir = tmp_path / "ref", tmp_path / "got"
    ref_dir.mkdir(); got_dir.mkdir()
    names = [f"img_{b}.png" for b in range(B)]
    masks = [f"mask_{b}.png" for b in range(B)]
    for b in range(B):                                                            # the reference's procedure
        cv2.imwrite(str(ref_dir / names[b]), x_rgba[b].numpy())
        cv2.imwrite(str(ref_dir / masks[b]), x[b].numpy())
        cv2.imwrite(str(ref_dir / names[b].replace(".png", "_ori.png")), ori[b].numpy())
    with nb.AttackImageSink(cuda) as sink:
        sink.put(x_rgba.to(cuda), x.to(cuda), ori.to(cuda), [str(got_dir / n) for n in names], [str(got_dir / m) for m in masks])
    files = sorted(os.listdir(ref_dir))
    assert files == sorted(os.listdir(got_dir)) and len(files) == 3 * B
    for f in files:
        assert (ref_dir / f).read_bytes() == (got_dir / f).read_bytes(), f


def test_philox_kernel_matches_the_published_algorithm_and_feeds_the_sampler(cuda):
    """The kernel-side generator is bit-for-bit Philox4x32-10 (numpy restatement checked against Random123's known-answer
    vectors on the CPU); nfb_coarse_z_rng / nfb_hierarchical_rng consume exactly those numbers: their outputs equal the
    explicit-t_rand / explicit-u entry points fed with nfb_philox_uniform's streams 0 / 1."""
    from nerfail_b200 import ops
    seed, offset = 0x1234_5678_9ABC_DEF0, 7
    for sid in (0, 1):
        got = ops.philox_uniform(seed, offset, sid, 4099, cuda).cpu().numpy()
        assert np.array_equal(got, philox_uniform_ref(seed, offset, sid, 4099))
    R, S, N = 777, 64, 128
    K, _ = synth.intrinsics(40, 40)
    rays = no.camera_rays(40, 40, K, torch.tensor(synth.pose_spherical(10.0, -30.0, 4.0)[:3, :4]), 2.0, 6.0)[:R].to(cuda)
    t = ops.philox_uniform(seed, offset, 0, R * S, cuda).reshape(R, S)
    z_a, z_b = ops.coarse_z(rays, S, False, None, rng=(seed, offset)), ops.coarse_z(rays, S, False, t)
    assert torch.equal(z_a, z_b)
    assert bool((z_a[:, 1:] >= z_a[:, :-1]).all()) and float(z_a.min()) >= 2.0 and float(z_a.max()) <= 6.0
    w = torch.rand(R, S, device=cuda) ** 3
    u = ops.philox_uniform(seed, offset, 1, R * N, cuda).reshape(R, N)
    f_a, f_b = ops.hierarchical(z_a, w, N, None, rng=(seed, offset)), ops.hierarchical(z_a, w, N, u)
    for a, b in zip(f_a, f_b):
        assert torch.equal(a, b)
    # uniformity of a long stream: mean 1/2, variance 1/12, no short-range correlation, all 24-bit values in [0, 1)
    x = ops.philox_uniform(99, 1, 0, 1 << 22, cuda).double()
    assert abs(float(x.mean()) - 0.5) < 1e-3 and abs(float(x.var()) - 1 / 12) < 1e-3
    assert abs(float(((x[1:] - 0.5) * (x[:-1] - 0.5)).mean())) < 2e-4
    assert float(x.min()) >= 0.0 and float(x.max()) < 1.0
    hist = torch.histc(x.float(), bins=64, min=0.0, max=1.0)
    assert float((hist - x.numel() / 64).abs().max()) < 5 * (x.numel() / 64) ** 0.5


def test_stochastic_render_is_reproducible_and_statistically_equal_to_torch_draws(cuda, monkeypatch):
    """render(perturb = 1) on the fused path draws its stratified samples in the kernels: torch.manual_seed makes it
    reproducible, a different seed changes the image, and the image agrees with the one rendered from torch.rand draws
    (NERFAIL_B200_RNG=torch) to within the sampling noise of either."""
    import nerfail_b200 as nb
    _, kw = make_kwargs(cuda)
    kw = dict(kw, perturb=1.0)
    H = W = 48
    K, _ = synth.intrinsics(H, W)
    c2w = torch.tensor(synth.pose_spherical(30.0, -30.0, 4.0)[:3, :4])
    with torch.no_grad():
        def once(seed):
            torch.manual_seed(seed)
            return nb.render(H, W, K, chunk=1024, c2w=c2w, near=2., far=6., **kw)[0]
        a, b, c = once(3), once(3), once(4)
        det = nb.render(H, W, K, chunk=1024, c2w=c2w, near=2., far=6., **dict(kw, perturb=0.0))[0]
        monkeypatch.setenv("NERFAIL_B200_RNG", "torch")
        t1, t2 = once(3), once(4)
    assert torch.equal(a, b) and not torch.equal(a, c)
    noise_torch = float((t1 - t2).abs().mean())                 # sampling noise between two torch-drawn renders
    assert noise_torch > 0
    assert float((a - c).abs().mean()) < 2.0 * noise_torch and float((a - c).abs().mean()) > 0.5 * noise_torch
    assert float((a - t1).abs().mean()) < 2.0 * noise_torch
    assert float((a - det).abs().mean()) < 3.0 * noise_torch
