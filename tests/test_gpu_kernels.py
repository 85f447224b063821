"""GPU parity of the HBM-bound kernels (rays, depths, compositing, sampling, GaussNet, 8-NN) through the C ABI,
against the oracle and the golden vectors minted from the reference.  Tolerances are stated per test."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import gauss_oracle as go
from oracle import nerf_oracle as no
from oracle import synth

pytestmark = pytest.mark.gpu
T = torch.from_numpy


def close(a, b, rtol, atol, what=""):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, equal_nan=True, err_msg=what)


def test_device_is_blackwell(cuda):
    from nerfail_b200 import _lib
    assert _lib.load().nfb_device_cc() // 10 == 10


def test_get_rays_bit_exact(cuda):
    from nerfail_b200 import ops
    for H, W, th in ((12, 12, 30.0), (37, 53, -110.0)):
        K, _ = synth.intrinsics(H, W)
        c2w = synth.pose_spherical(th, -30.0, 4.0)[:3, :4]
        got = ops.get_ray_batch(H, W, K, c2w, 2.0, 6.0, device=cuda).cpu()
        want = no.camera_rays(H, W, K, torch.tensor(c2w), 2.0, 6.0)
        # same fp32 operations in the same order, no contraction: expected identical; allow 1 ulp on the normalisation
        close(got[:, :8], want[:, :8], 0, 0, "origins / directions / bounds")
        close(got[:, 8:], want[:, 8:], 3e-7, 0, "unit view directions (1 ulp: sqrt/div rounding)")


def test_coarse_depths(cuda):
    from nerfail_b200 import ops
    rays = torch.zeros(50, 11); rays[:, 6] = 2.0; rays[:, 7] = 6.0; rays[10:, 6] = 0.5; rays[10:, 7] = 9.25
    g = torch.Generator().manual_seed(0)
    t_rand = torch.rand(50, 64, generator=g)
    for lindisp in (False, True):
        for tr in (None, t_rand):
            got = ops.coarse_z(rays.to(cuda), 64, lindisp, None if tr is None else tr.to(cuda)).cpu()
            want = no.coarse_depths(rays, 64, lindisp, tr)
            close(got, want, 3e-7, 0, f"lindisp={lindisp} jitter={tr is not None}")   # <= 2 ulp (division in lindisp)
    # ragged sample counts
    for S in (1, 2, 7, 33):
        close(ops.coarse_z(rays.to(cuda), S).cpu(), no.coarse_depths(rays, S), 3e-7, 0)


def test_composite_forward_vs_reference_golden(cuda):
    from nerfail_b200 import ops
    g = golden("composite.npz")
    raw, z, rd = T(g["raw"]).to(cuda), T(g["z"]).to(cuda), T(g["rays_d"]).to(cuda)
    for white in (False, True):
        rgb, disp, acc, w, depth = ops.composite_fwd(raw, z, rd, None, white)
        # fp32 with a warp-tree product/sum order instead of the sequential one: tolerance 1e-5 relative; the absolute
        # term covers alpha = 1 - exp(-x) for tiny x, where one ulp of exp() is 6e-8 of alpha
        close(w, g[f"weights_w{int(white)}"], 1e-5, 2e-7, "weights")
        close(rgb, g[f"rgb_w{int(white)}"], 1e-5, 1e-6, "rgb")
        close(acc, g[f"acc_w{int(white)}"], 1e-5, 1e-6, "acc")
        close(depth, g[f"depth_w{int(white)}"], 1e-5, 1e-6, "depth")
        close(disp, g[f"disp_w{int(white)}"], 1e-4, 1e-6, "disp (NaN where acc == 0, like the reference)")


def test_composite_backward_vs_reference_autograd(cuda):
    from nerfail_b200 import ops
    g = golden("composite.npz")
    raw, z, rd = T(g["raw"])[8:].to(cuda), T(g["z"])[8:].to(cuda), T(g["rays_d"])[8:].to(cuda)
    raw = raw.clone().requires_grad_(True)
    outs = ops.CompositeFn.apply(raw, z, rd, None, True)
    sum((o * T(g[f"cot{i}"]).to(cuda)).sum() for i, o in enumerate(outs)).backward()
    ref = g["g_raw"]
    err = np.abs(raw.grad.cpu().numpy() - ref)
    # north_star tolerance for input gradients: 1e-3 relative (to the gradient scale of the ray)
    scale = np.abs(ref).max(axis=(1, 2), keepdims=True) + 1e-12
    assert (err / scale).max() < 1e-3, (err / scale).max()


def test_composite_edge_shapes(cuda):
    from nerfail_b200 import ops
    g = torch.Generator().manual_seed(5)
    for R, S in ((1, 1), (3, 5), (7, 33), (2, 192), (5, 300)):
        raw = torch.randn(R, S, 4, generator=g)
        z = torch.sort(torch.rand(R, S, generator=g) * 4 + 2, -1).values
        rd = torch.randn(R, 3, generator=g)
        got = ops.composite_fwd(raw.to(cuda), z.to(cuda), rd.to(cuda), None, True)
        want = no.composite(raw, z, rd, True)
        for a, b in zip(got, want):
            close(a, b, 2e-5, 1e-6, f"R={R} S={S}")
        rg = raw.clone().requires_grad_(True)
        wo = no.composite(rg, z, rd, True)
        (wo[0].sum() + 0.3 * wo[4].sum() + 0.1 * wo[2].sum()).backward()
        rc = raw.to(cuda).requires_grad_(True)
        co = ops.CompositeFn.apply(rc, z.to(cuda), rd.to(cuda), None, True)
        (co[0].sum() + 0.3 * co[4].sum() + 0.1 * co[2].sum()).backward()
        close(rc.grad, rg.grad, 1e-3, 1e-6, f"grad R={R} S={S}")
    # empty batch
    out = ops.composite_fwd(torch.zeros(0, 8, 4, device=cuda), torch.zeros(0, 8, device=cuda), torch.zeros(0, 3, device=cuda))
    assert out[0].shape == (0, 3)


def test_pts_max_first_maximum(cuda):
    from nerfail_b200 import ops
    g = torch.Generator().manual_seed(9)
    R, S = 40, 192
    raw = torch.randn(R, S, 4, generator=g) * 3
    raw[:5, :, 3] = -1.0                                      # all weights zero -> argmax = 0 (first maximum)
    z = torch.sort(torch.rand(R, S, generator=g) * 4 + 2, -1).values
    rays = torch.randn(R, 11, generator=g)
    got = ops.composite_fwd(raw.to(cuda), z.to(cuda), rays.to(cuda), None, True, want_pts_max=True)
    w = got[3].cpu()
    best = torch.argmax(w, dim=1)                             # nerf_to_coord.py:418 on the kernel's own weights
    want = rays[:, 0:3] + rays[:, 3:6] * z[torch.arange(R), best][:, None]
    close(got[5], want, 0, 0, "pts_max must be bit-exact: o + d * z[argmax]")
    assert (best[:5] == 0).all()


def test_sample_pdf_vs_reference_golden(cuda):
    from nerfail_b200 import ops
    g = golden("sample_pdf.npz")
    bins, w = T(g["bins"]), T(g["weights"])
    det, inds = ops.sample_pdf(bins.to(cuda), w.to(cuda), 128, None, return_inds=True)
    _, inds_ref = no.inverse_cdf_samples(bins, w, 128, return_inds=True)
    rnd, inds_r = ops.sample_pdf(bins.to(cuda), w.to(cuda), 128, T(g["u_rnd"]).to(cuda), return_inds=True)
    _, inds_r_ref = no.inverse_cdf_samples(bins, w, 128, T(g["u_rnd"]), return_inds=True)
    # sample indexing: the searchsorted results must be identical wherever u is not within float rounding of a CDF
    # knot (the pdf normaliser is summed in a different order than ATen's vectorised CPU sum, so CDF values may
    # differ in the last bit); at such knife edges either neighbouring bin yields the same sample.
    wn = w + 1e-5
    cdf = torch.cat([torch.zeros(w.shape[0], 1), torch.cumsum(wn / wn.sum(-1, keepdim=True), -1)], -1)
    for got_i, ref_i, u in ((inds, inds_ref, torch.linspace(0, 1, 128).expand(w.shape[0], 128)),
                            (inds_r, inds_r_ref, T(g["u_rnd"]))):
        edge = ((u[:, :, None] - cdf[:, None, :]).abs().min(-1).values < 1e-6)
        bad = (got_i.cpu().long() != ref_i) & ~edge
        assert int(bad.sum()) == 0, f"{int(bad.sum())} index mismatches away from CDF knots"
        assert float(edge.float().mean()) < 0.05
    close(det, g["det"], 1e-5, 1e-5, "deterministic samples")
    close(rnd, g["rnd"], 1e-5, 1e-5, "random-u samples")


def test_sample_indexing_bit_exact_on_exact_cdf(cuda):
    """With weights whose normalised CDF is exactly representable (powers of two) every summation order gives
    the same CDF, so the searchsorted indices must match the reference bit for bit."""
    from nerfail_b200 import ops
    R, nb = 16, 65
    w = torch.full((R, nb - 1), 1.0) - 1e-5            # + 1e-5 inside sample_pdf -> exactly 1.0 each, sum = 64
    bins = torch.linspace(2, 6, nb).expand(R, nb).contiguous()
    g = torch.Generator().manual_seed(3)
    u = torch.rand(R, 128, generator=g)
    u[:, :8] = torch.tensor([0.0, 1.0, 0.5, 0.25, 1 / 64, 63 / 64, 0.999999, 1e-9])   # knots and ends
    s, inds = ops.sample_pdf(bins.to(cuda), w.to(cuda), 128, u.to(cuda), return_inds=True)
    s_ref, inds_ref = no.inverse_cdf_samples(bins, w, 128, u, return_inds=True)
    assert torch.equal(inds.cpu().long(), inds_ref)
    close(s, s_ref, 1e-6, 1e-6)


def test_hierarchical_merge(cuda):
    from nerfail_b200 import ops
    g = torch.Generator().manual_seed(21)
    for R, Sc, N, random_u in ((9, 64, 128, False), (9, 64, 128, True), (4, 16, 40, True), (3, 3, 5, False)):
        zc = torch.sort(torch.rand(R, Sc, generator=g) * 4 + 2, -1).values
        w = torch.rand(R, Sc, generator=g) ** 3
        u = torch.rand(R, N, generator=g) if random_u else None
        zf, zs, zstd = ops.hierarchical(zc.to(cuda), w.to(cuda), N, None if u is None else u.to(cuda))
        zf_ref, zs_ref, zstd_ref = no.hierarchical_depths(zc, w, N, u)
        close(zs, zs_ref, 1e-5, 1e-5, "new depths")
        close(zf, zf_ref, 1e-5, 1e-5, "merged depths")
        close(zstd, zstd_ref, 1e-4, 1e-6, "z_std")
        zf_c = zf.cpu()
        assert (zf_c[:, 1:] >= zf_c[:, :-1]).all(), "merged depths must be sorted"
        # merge is a permutation of cat(coarse, new): bit-exact multiset equality
        assert torch.equal(torch.sort(torch.cat([zc, zs.cpu()], -1), -1).values, zf_c)


def test_positional_encoding_to_hbm(cuda):
    import nerfail_b200 as nb
    g = golden("mlp.npz")
    x = T(g["x"]).to(cuda)
    e10, n10 = nb.get_embedder(10)
    e4, n4 = nb.get_embedder(4)
    assert (n10, n4) == (63, 27)
    # sinf/cosf vs the CPU's vectorised sin/cos: <= 2 ulp of values in [-1,1]
    close(e10(x), g["enc10"], 0, 3e-7)
    close(e4(x), g["enc4"], 0, 3e-7)


def test_gauss_weights_vs_reference_golden(cuda):
    import nerfail_b200 as nb
    g = golden("gauss.npz")
    i_w, dist = nb.create_gauss_w(cuda, 0.02)(T(g["dist_idx"]).to(cuda))
    close(i_w[:, 1], g["i_w"][:, 1], 0, 0, "indices are copied")
    close(i_w[:, 0], g["i_w"][:, 0], 2e-6, 1e-7, "weights")       # expf vs the CPU's exp: few ulp
    assert dist.shape == (g["dist_idx"].shape[0], 1) + g["dist_idx"].shape[2:]


def test_gauss_gather_and_scatter_vs_reference_golden(cuda):
    import nerfail_b200 as nb
    g = golden("gauss.npz")
    s, i_w, ori = T(g["spatial_rgb"]).to(cuda), T(g["i_w"]).to(cuda), T(g["ori"]).to(cuda)
    for tag, eps in (("none", None), ("e32", 32), ("e2", 2)):
        net = nb.gauss_net(cuda, 0.02, None, "my_model", epsilon=eps)
        sg = s.clone().requires_grad_(True)
        x, x_rgba = net.perturbed(sg, i_w, ori)
        close(x, g[f"x_{tag}"], 1e-5, 1e-5, f"x eps={eps}")
        close(x_rgba, g[f"xrgba_{tag}"], 1e-5, 1e-4, f"x_rgba eps={eps}")
        assert abs(net.epsilon_3d_min - float(g[f"epsmin_{tag}"])) < 1e-3
        assert abs(net.epsilon_3d_max - float(g[f"epsmax_{tag}"])) < 1e-3
        ((x * T(g[f"cx_{tag}"]).to(cuda)).sum() + (x_rgba * T(g[f"cr_{tag}"]).to(cuda)).sum()).backward()
        ref = g[f"grad_{tag}"]
        err = np.abs(sg.grad.cpu().numpy() - ref).max()
        assert err < 1e-3 * np.abs(ref).max(), (tag, err, np.abs(ref).max())     # 1e-3 relative (atomics reorder sums)


def test_gauss_full_forward_signature_and_double_backward(cuda):
    import nerfail_b200 as nb

    class Head(torch.nn.Module):
        def forward(self, x):
            return x.mean(dim=(2, 3)).repeat(1, 3)[:, :8]
    s, di, ori = synth.gauss_inputs(1, P=3, H=32, W=32, B=2)
    i_w, _ = nb.create_gauss_w(cuda, 0.02)(di.to(cuda))
    net = nb.gauss_net(cuda, 0.02, Head(), "my_model", epsilon=32)
    sg = s.to(cuda).requires_grad_(True)
    x, x_rgba, cla, ori_f, ori_cla = net(sg, i_w, ori.to(cuda), False)
    assert x.shape == (2, 32, 32, 4) and cla.shape == (2, 8) and ori_f.dtype == torch.float32
    # deepfool.py:76-77 asks for create_graph=True; the graph must build and be differentiable again
    g1 = torch.autograd.grad(cla[:, 1].sum(), sg, retain_graph=True, create_graph=True)[0]
    assert g1.shape == sg.shape and torch.isfinite(g1).all()
    # oracle comparison of the first-order gradient
    so = s.clone().requires_grad_(True)
    xo, xro, _ = go.gauss_forward(so, go.gaussian_weights(di, 0.02), ori, 32)
    chw = xro.permute(0, 3, 1, 2)
    cla_o = Head()(torch.where(chw[:, 3:4] > 0, chw[:, :3], torch.full_like(chw[:, :3], 255.0)))
    g_ref = torch.autograd.grad(cla_o[:, 1].sum(), so)[0]
    close(g1, g_ref, 1e-3, 1e-7 + 1e-3 * float(g_ref.abs().max()))
    # second order: d/ds <g1, v> exists (zero almost everywhere except through alpha*x products)
    v = torch.randn_like(g1)
    g2 = torch.autograd.grad((g1 * v).sum(), sg, allow_unused=True)[0]
    assert g2 is None or torch.isfinite(g2).all()


def test_classifier_input_conversion_and_its_derivatives(cuda):
    """nfb_rgba_to_chw / nfb_chw_to_rgba (model/GaussNet.py:121-145: NHWC RGBA -> NCHW RGB, 255 where alpha is 0) against the
    reference's torch expression: values bit-exact (float and uint8 image), first and second derivative exact."""
    from nerfail_b200 import ops
    g = torch.Generator().manual_seed(3)
    img = (torch.rand(2, 13, 7, 4, generator=g) * 255)
    img[..., 3] = torch.where(torch.rand(2, 13, 7, generator=g) < 0.4, torch.zeros(()), img[..., 3])

    def ref(t):
        chw = t.permute(0, 3, 1, 2)
        return torch.where(chw[:, 3:4] > 0, chw[:, :3], torch.full_like(chw[:, :3], 255.0))
    a = img.clone().to(cuda).requires_grad_(True)
    b = img.clone().requires_grad_(True)
    ya, yb = ops.RgbaToChwFn.apply(a, 255.0), ref(b)
    assert torch.equal(ya.detach().cpu(), yb.detach())
    u8 = img.to(torch.uint8)
    assert torch.equal(ops.rgba_u8_to_chw(u8.to(cuda)).cpu(), ref(u8.float()))
    w = torch.randn(yb.shape, generator=g)
    ga = torch.autograd.grad((ya * w.to(cuda)).sum(), a, create_graph=True)[0]
    gb = torch.autograd.grad((yb * w).sum(), b, create_graph=True)[0]
    assert torch.equal(ga.detach().cpu(), gb.detach())
    # the gradient is linear in the upstream gradient: differentiate <ga, v> with respect to it
    wa = w.clone().to(cuda).requires_grad_(True)
    wb = w.clone().requires_grad_(True)
    v = torch.randn(img.shape, generator=g)
    ga2 = torch.autograd.grad((ops.RgbaToChwFn.apply(a, 255.0) * wa).sum(), a, create_graph=True)[0]
    gb2 = torch.autograd.grad((ref(b) * wb).sum(), b, create_graph=True)[0]
    ha = torch.autograd.grad((ga2 * v.to(cuda)).sum(), wa)[0]
    hb = torch.autograd.grad((gb2 * v).sum(), wb)[0]
    assert torch.equal(ha.cpu(), hb)


def test_attack_sign_step_on_active_rows_equals_full_table_update(cuda):
    """dist.attack_sign_step_ (attack_NeRFail_S.py:357-392) with the active-row exchange (nfb_attack_pack_rgb +
    nfb_attack_sign_step: only the RGB of rows with A > 0 is packed, reduced and updated) against the full-table torch
    expression, descending and ascending, including rows at the clamp and zero gradients."""
    from nerfail_b200 import dist as nd
    g = torch.Generator().manual_seed(9)
    P, H, W = 2, 17, 13
    init = torch.randn(P, H, W, 4, generator=g) * 3
    init[..., 3] = (torch.rand(P, H, W, generator=g) > 0.45).float() * 255
    cur = init.clone()
    cur[..., :3] += (torch.rand(P, H, W, 3, generator=g) - 0.5) * 6.0          # some rows sit at init +- eps already
    cur[..., :3] = torch.max(torch.min(cur[..., :3], init[..., :3] + 2.5), init[..., :3] - 2.5)
    grad = torch.randn(P, H, W, 4, generator=g)
    grad[0, :3] = 0.0                                                          # sign(0) = 0: no step
    for minimise in (True, False):
        want = nd.attack_sign_step_(cur.clone(), grad.clone(), init, 2.0, 2.5, minimise=minimise)
        s = cur.clone().to(cuda)
        idx = nd.active_rows(s)
        got = nd.attack_sign_step_(s, grad.clone().to(cuda), init.to(cuda), 2.0, 2.5, minimise=minimise, active_idx=idx)
        assert torch.equal(got.cpu(), want)
        packed = nd.allreduce_active_rgb(grad.to(cuda), idx)
        assert torch.equal(packed.cpu(), grad.reshape(-1, 4)[idx.cpu(), :3])


def test_knn8_bit_exact_indices(cuda):
    from nerfail_b200 import ops
    g = golden("knn.npz")
    q, c = g["query"].reshape(-1, 3), g["cand"]
    d_ref, i_ref = go.knn8_exact(q, c)
    d, i = ops.knn8(T(q).to(cuda), T(c).to(cuda))
    assert np.array_equal(i.cpu().numpy(), i_ref), "8-NN indices must be bit-exact against the direct-difference oracle"
    assert np.array_equal(d.cpu().numpy(), d_ref), "distances are sqrt of the same fp32 d2"
    # reference on-disk layout [2,H,W,8] float32 with indices as floats
    packed = ops.knn8_dist_idx(T(g["query"]).to(cuda), T(c).to(cuda)).cpu().numpy()
    assert packed.shape == (2, 10, 16, 8) and packed.dtype == np.float32
    assert np.array_equal(packed[1].reshape(-1, 8).astype(np.int32), i_ref)
    # statistics against the reference's own cdist procedure: report, do not gate tightly
    agree_direct = (i.cpu().numpy() == g["i_ref_direct"].reshape(-1, 8)).all(1).mean()
    assert agree_direct > 0.99


def test_knn8_ties_duplicates_and_ragged(cuda):
    from nerfail_b200 import ops
    rng = np.random.default_rng(4)
    cand = rng.integers(0, 4, size=(5000, 3)).astype(np.float32)      # many exact ties on an integer lattice
    qry = rng.integers(0, 4, size=(301, 3)).astype(np.float32)
    d_ref, i_ref = go.knn8_exact(qry, cand)
    d, i = ops.knn8(T(qry).to(cuda), T(cand).to(cuda))
    assert np.array_equal(i.cpu().numpy(), i_ref), "ties must resolve to the lowest candidate index"
    assert np.array_equal(d.cpu().numpy(), d_ref)
    # exactly 8 candidates, one query
    d, i = ops.knn8(T(qry[:1]).to(cuda), T(cand[:8]).to(cuda))
    assert sorted(i.cpu().numpy()[0].tolist()) == list(range(8))


def test_knn_grid_equals_brute_force_bit_for_bit(cuda):
    """The grid-accelerated search must return exactly the brute-force result (indices AND distances): golden set and
    tie lattice against the oracle; a 150 k-point noisy sphere surface with outliers against the brute-force kernel
    (itself pinned to the oracle above), with queries on the surface, far outside the bounding box and at candidates."""
    from nerfail_b200 import ops
    g = golden("knn.npz")
    q, c = g["query"].reshape(-1, 3), g["cand"]
    d_ref, i_ref = go.knn8_exact(q, c)
    grid = ops.KnnGrid(T(c).to(cuda))
    d, i = grid.query(T(q).to(cuda))
    assert np.array_equal(i.cpu().numpy(), i_ref) and np.array_equal(d.cpu().numpy(), d_ref)
    packed = grid.query_dist_idx(T(g["query"]).to(cuda)).cpu().numpy()
    assert packed.shape == (2, 10, 16, 8) and np.array_equal(packed[1].reshape(-1, 8).astype(np.int32), i_ref)

    rng = np.random.default_rng(4)
    cand = rng.integers(0, 4, size=(5000, 3)).astype(np.float32)      # exact ties and ~78-fold duplicates per lattice site
    qry = rng.integers(0, 4, size=(301, 3)).astype(np.float32)
    d_ref, i_ref = go.knn8_exact(qry, cand)
    d, i = ops.KnnGrid(T(cand).to(cuda)).query(T(qry).to(cuda))
    assert np.array_equal(i.cpu().numpy(), i_ref), "ties must resolve to the lowest candidate index"
    assert np.array_equal(d.cpu().numpy(), d_ref)

    n = 150_000
    dirs = rng.normal(size=(n, 3)); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    cand = (dirs * (1.0 + 0.002 * rng.normal(size=(n, 1)))).astype(np.float32)
    cand[:200] = rng.uniform(-6, 6, size=(200, 3)).astype(np.float32)          # background-like outliers stretch the box
    cand[200:260] = cand[260:320]                                                # duplicates with different indices
    qd = rng.normal(size=(30_000, 3)); qd /= np.linalg.norm(qd, axis=1, keepdims=True)
    qry = (qd * (1.0 + 0.002 * rng.normal(size=(30_000, 1)))).astype(np.float32)
    qry[:500] = rng.uniform(-20, 20, size=(500, 3)).astype(np.float32)           # far outside the candidates' bounding box
    qry[500:1000] = cand[1000:1500]                                              # zero distances
    qry[1000:1500] = rng.uniform(-6, 6, size=(500, 3)).astype(np.float32)        # empty space inside the box
    cg, qg = T(cand).to(cuda), T(qry).to(cuda)
    d_bf, i_bf = ops.knn8(qg, cg)
    grid = ops.KnnGrid(cg)
    stats = torch.zeros(1, dtype=torch.int64, device=cuda)
    d, i = grid.query(qg, stats)
    assert torch.equal(i, i_bf), f"{int((i != i_bf).any(1).sum())} queries differ from the brute-force search"
    assert torch.equal(d, d_bf)
    d_ref, i_ref = go.knn8_exact(qry[:1600:8], cand)
    assert np.array_equal(i.cpu().numpy()[:1600:8], i_ref) and np.array_equal(d.cpu().numpy()[:1600:8], d_ref)
    pruning = qry.shape[0] * n / float(stats.item())
    print(f"grid kNN pruning factor {pruning:.0f}x ({float(stats.item()) / qry.shape[0]:.0f} distance evaluations per query)")
    assert pruning > 20


def test_fused_adam_matches_torch_adam_and_shares_its_state_dict(cuda):
    """nfb_adam_step against torch.optim.Adam on the CPU (the reference's optimizer, run_nerf.py:213/:792) over several
    steps with a decaying learning rate (:796-800); then the two optimizers swap state_dicts (checkpoint compatibility,
    :228) and keep agreeing.  Tolerance: ulp-level differences of the fused fp32 update (2e-6 relative on the parameters)."""
    import nerfail_b200 as nb
    g = torch.Generator().manual_seed(3)
    shapes = [(256, 63), (256,), (1, 256), (1,), (3, 128), (128, 283), (4097,)]
    ref_p = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in shapes]
    got_p = [torch.nn.Parameter(p.detach().clone().to(cuda)) for p in ref_p]
    ref_o = torch.optim.Adam(ref_p, lr=5e-4, betas=(0.9, 0.999))
    got_o = nb.Adam(got_p, lr=5e-4, betas=(0.9, 0.999))

    def one_step(i):
        for a, b in zip(ref_p, got_p):
            a.grad = torch.randn(a.shape, generator=g) * (10.0 ** float(torch.randint(-4, 2, (1,), generator=g)))
            b.grad = a.grad.to(cuda)
        v0 = [p._version for p in got_p]
        ref_o.step(); got_o.step()
        assert all(p._version > v for p, v in zip(got_p, v0)), "the fused step must bump tensor versions (weight re-pack trigger)"
        lr = nb.decayed_lrate(5e-4, 250, i + 1)
        nb.set_lrate(ref_o, lr); nb.set_lrate(got_o, lr)

    def check():
        for a, b in zip(ref_p, got_p):
            assert torch.allclose(b.detach().cpu(), a.detach(), rtol=2e-6, atol=2e-7), float((b.detach().cpu() - a.detach()).abs().max())
        for a, b in zip(ref_p, got_p):
            sa, sb = ref_o.state[a], got_o.state[b]
            assert float(sa["step"]) == float(sb["step"])
            assert torch.allclose(sb["exp_avg"].cpu(), sa["exp_avg"], rtol=1e-5, atol=1e-6 * float(sa["exp_avg"].abs().max()))
            assert torch.allclose(sb["exp_avg_sq"].cpu(), sa["exp_avg_sq"], rtol=1e-5, atol=1e-6 * float(sa["exp_avg_sq"].abs().max()))

    for i in range(4):
        one_step(i)
    check()
    # swap checkpoints: torch state into the fused optimizer and the fused state into torch
    sd_ref, sd_got = ref_o.state_dict(), got_o.state_dict()
    assert sd_ref["param_groups"][0].keys() == sd_got["param_groups"][0].keys()
    got_o.load_state_dict(sd_ref)
    ref_o.load_state_dict({"state": {k: {kk: (vv.cpu() if torch.is_tensor(vv) else vv) for kk, vv in v.items()} for k, v in sd_got["state"].items()},
                           "param_groups": sd_got["param_groups"]})
    for i in range(4, 7):
        one_step(i)
    check()
    # parameters without a gradient are skipped, like torch
    got_p[0].grad = None
    before = got_p[0].detach().clone()
    got_o.step()
    assert torch.equal(got_p[0].detach(), before)


def test_rays_from_batch_and_fused_mse_loss(cuda):
    """The two small fusions of the training step: the [N,11] ray batch of render(rays=batch_rays) (run_nerf.py:95-123) bit
    for bit against the oracle's torch restatement, and img2mse(fine) + img2mse(coarse) with its gradient (run_nerf.py:781-791)
    against torch autograd."""
    from nerfail_b200 import ops
    g = torch.Generator().manual_seed(12)
    for n in (1, 513, 4096):
        o, d = torch.randn(n, 3, generator=g), torch.randn(n, 3, generator=g) * 3
        want = no.rays_from_batch(torch.stack([o, d], 0), 2.0, 6.0)
        got = ops.rays_from_batch(o.to(cuda), d.to(cuda), 2.0, 6.0).cpu()
        assert torch.equal(got[:, :8], want[:, :8])
        assert float((got[:, 8:] - want[:, 8:]).abs().max()) <= 1.2e-7          # d / |d|: torch.norm may round the sum differently by 1 ulp
        rgb, rgb0, tgt = torch.rand(n, 3, generator=g), torch.rand(n, 3, generator=g), torch.rand(n, 3, generator=g)
        a, a0 = rgb.clone().requires_grad_(True), rgb0.clone().requires_grad_(True)
        ref = torch.mean((a - tgt) ** 2) + torch.mean((a0 - tgt) ** 2)
        ref.backward()
        b, b0 = rgb.to(cuda).requires_grad_(True), rgb0.to(cuda).requires_grad_(True)
        loss, mse, psnr = ops.MseLoss2Fn.apply(b, b0, tgt.to(cuda))
        assert abs(float(psnr[0]) + 10 * np.log10(float(mse[0]))) < 1e-4 and abs(float(psnr[1]) + 10 * np.log10(float(mse[1]))) < 1e-4
        (loss * 1.0).backward()
        assert abs(float(loss) - float(ref)) <= 2e-6 * float(ref) + 1e-9
        assert abs(float(mse[0]) - float(torch.mean((rgb - tgt) ** 2))) <= 2e-6 and abs(float(mse[1]) - float(torch.mean((rgb0 - tgt) ** 2))) <= 2e-6
        assert float((b.grad.cpu() - a.grad).abs().max()) <= 1e-6 * float(a.grad.abs().max()) + 1e-12
        assert float((b0.grad.cpu() - a0.grad).abs().max()) <= 1e-6 * float(a0.grad.abs().max()) + 1e-12
        c = rgb.to(cuda).requires_grad_(True)                                    # no coarse image (N_importance = 0)
        l1, m1, _ = ops.MseLoss2Fn.apply(c, None, tgt.to(cuda))
        l1.backward()
        assert abs(float(l1) - float(torch.mean((rgb - tgt) ** 2))) <= 2e-6 and float(m1[1]) == 0.0


def test_hierarchical_fixed_shape_kernel_equals_the_general_kernel(cuda, monkeypatch):
    """The compile-time-sized kernel for the lego shape (64 coarse depths, 128 new samples; csrc/sampling.cu) against the
    general one (NERFAIL_B200_HIER=generic), bit for bit: deterministic u, caller-provided random u, kernel-side Philox u;
    flat, peaked and mostly-empty weight profiles (long runs of empty bins are where the search shortcuts could go wrong)."""
    from nerfail_b200 import ops
    g = torch.Generator().manual_seed(31)
    R = 1500
    z = torch.sort(torch.rand(R, 64, generator=g) * 4 + 2, dim=-1).values
    w = torch.rand(R, 64, generator=g) ** 6
    w[:200] = 0.0                                            # flat pdf
    w[200:400, 10:] = 0.0                                     # mass in the first bins only
    w[400:600] = 0.0; w[400:600, 40] = 1.0                    # a single occupied bin
    w[600:700] = torch.rand(100, 64, generator=g)             # broad
    z[700:720] = torch.linspace(2, 6, 64)                     # the reference's own deterministic coarse depths
    u = torch.rand(R, 128, generator=g)
    zc, wc, uc = z.to(cuda), w.to(cuda), u.to(cuda)
    for kwargs in (dict(u=None), dict(u=uc), dict(u=None, rng=(123, 9))):
        monkeypatch.delenv("NERFAIL_B200_HIER", raising=False)
        a = ops.hierarchical(zc, wc, 128, **kwargs)
        monkeypatch.setenv("NERFAIL_B200_HIER", "generic")
        b = ops.hierarchical(zc, wc, 128, **kwargs)
        for x, y in zip(a, b):
            assert torch.equal(x, y), kwargs
        assert bool((a[0][:, 1:] >= a[0][:, :-1]).all())
