"""Tensor-level wrappers over the C ABI and the autograd.Functions built from them.

Every function here launches hand-written sm_100a kernels through ctypes; PyTorch only owns the memory
and the stream.  Reference citations (file:line, relative to the NeRFail tree) name the code each op replaces.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream


def _f32(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("nerfail_b200: CUDA tensors required (the B200 kernels have no CPU fallback)")


# ------------------------------------------------------------------------------------------------
# rays / depths
# ------------------------------------------------------------------------------------------------
def get_ray_batch(H: int, W: int, K, c2w, near: float, far: float, device=None) -> torch.Tensor:
    """[H*W, 11] ray batch (o, d, near, far, viewdir) for one pinhole camera.

    Replaces run_nerf_helpers.py:157-166 (get_rays) + run_nerf.py:102-123.
    """
    device = torch.device(device if device is not None else "cuda")
    K_h = np.ascontiguousarray(np.asarray(K, dtype=np.float64).reshape(3, 3))
    c2w_h = np.ascontiguousarray(
        (c2w.detach().cpu().numpy() if isinstance(c2w, torch.Tensor) else np.asarray(c2w)).astype(np.float32)[:3, :4])
    rays = torch.empty((H * W, 11), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        check(_lib.load().nfb_get_rays(H, W, K_h.ctypes.data, c2w_h.ctypes.data, float(near), float(far),
                                       ptr(rays), stream()), "nfb_get_rays")
    return rays


def rays_from_batch(rays_o: torch.Tensor, rays_d: torch.Tensor, near: float, far: float) -> torch.Tensor:
    """[N,11] ray batch (o, d, near, far, d / |d|) of render(rays=...) — run_nerf.py:95-123 for use_viewdirs, ndc=False."""
    _require_cuda(rays_o, rays_d)
    o, d = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3)
    rays = torch.empty((o.shape[0], 11), dtype=torch.float32, device=o.device)
    with torch.cuda.device(o.device):
        check(_lib.load().nfb_rays_from_batch(ptr(o), ptr(d), o.shape[0], float(near), float(far), ptr(rays), stream()),
              "nfb_rays_from_batch")
    return rays


class MseLoss2Fn(torch.autograd.Function):
    """loss = img2mse(rgb, target) + img2mse(rgb0, target) (run_nerf.py:781-789) with the gradient produced in the forward
    pass: one kernel instead of ~12 element-wise / reduction launches forward and ~8 backward.
    Returns (loss, mse [2], psnr [2]); mse and psnr (mse2psnr, run_nerf_helpers.py:10) are statistics, not differentiable."""

    @staticmethod
    def forward(ctx, rgb, rgb0, target):
        rgb, target = _f32(rgb), _f32(target)
        rgb0 = _f32(rgb0) if rgb0 is not None else None
        out = torch.empty(5, dtype=torch.float32, device=rgb.device)
        has0 = rgb0 is not None
        both = torch.empty((2 if has0 else 1,) + tuple(rgb.shape), dtype=torch.float32, device=rgb.device)   # g | g0: one buffer
        with torch.cuda.device(rgb.device):
            check(_lib.load().nfb_mse_loss2(ptr(rgb), ptr(rgb0), ptr(target), rgb.numel(), ptr(out), ptr(both[0]),
                                            ptr(both[1]) if has0 else None, stream()), "nfb_mse_loss2")
        ctx.save_for_backward(both)
        ctx.has0 = has0
        ctx.set_materialize_grads(False)           # no zero-filled gradients for the two statistics
        mse, psnr = out[1:3], out[3:5]
        ctx.mark_non_differentiable(mse, psnr)
        return out[0], mse, psnr

    @staticmethod
    def backward(ctx, g_loss, _g_mse, _g_psnr):
        if g_loss is None:
            return None, None, None
        scaled = ctx.saved_tensors[0] * g_loss     # both gradients in one launch
        return scaled[0], (scaled[1] if ctx.has0 else None), None


def coarse_z(rays: torch.Tensor, n_samples: int, lindisp: bool = False, t_rand: Optional[torch.Tensor] = None,
             rng: Optional[tuple] = None):
    """z_vals [R, n_samples] (run_nerf.py:357-379).  rng = (seed, offset): stratified jitter drawn in the kernel (Philox)."""
    _require_cuda(rays, t_rand)
    rays = _f32(rays)
    assert rays.shape[1] >= 8
    if rays.shape[1] != 11:
        padded = torch.zeros((rays.shape[0], 11), dtype=torch.float32, device=rays.device)
        padded[:, : rays.shape[1]] = rays
        rays = padded
    R = rays.shape[0]
    z = torch.empty((R, n_samples), dtype=torch.float32, device=rays.device)
    if t_rand is not None:
        t_rand = _f32(t_rand)
        assert t_rand.shape == (R, n_samples)
    with torch.cuda.device(rays.device):
        if rng is not None and t_rand is None:
            check(_lib.load().nfb_coarse_z_rng(ptr(rays), R, n_samples, int(bool(lindisp)), int(rng[0]), int(rng[1]), ptr(z),
                                               stream()), "nfb_coarse_z_rng")
        else:
            check(_lib.load().nfb_coarse_z(ptr(rays), R, n_samples, int(bool(lindisp)), ptr(t_rand), ptr(z), stream()),
                  "nfb_coarse_z")
    return z


# ------------------------------------------------------------------------------------------------
# compositing
# ------------------------------------------------------------------------------------------------
def _rays_d_view(rays_or_d: torch.Tensor):
    """Returns (tensor, pointer to d, pitch, has_origin)."""
    t = _f32(rays_or_d)
    if t.shape[-1] == 3:
        return t, t.data_ptr(), 3, False
    assert t.shape[-1] >= 6
    return t, t.data_ptr() + 3 * 4, t.shape[-1], True


def composite_fwd(raw, z_vals, rays_or_d, noise=None, white_bkgd=False, want_pts_max=False):
    """(rgb_map, disp, acc, weights, depth[, pts_max]) — run_nerf.py:262-305, nerf_to_coord.py:418-421."""
    _require_cuda(raw, z_vals, rays_or_d, noise)
    raw, z_vals = _f32(raw), _f32(z_vals)
    R, S = z_vals.shape
    assert raw.shape == (R, S, 4), f"raw {tuple(raw.shape)} vs z_vals {tuple(z_vals.shape)}"
    keep, dptr, pitch, has_o = _rays_d_view(rays_or_d)
    if want_pts_max and not has_o:
        raise RuntimeError("pts_max needs the full ray batch (origins), got directions only")
    dev = raw.device
    rgb = torch.empty((R, 3), dtype=torch.float32, device=dev)
    disp = torch.empty((R,), dtype=torch.float32, device=dev)
    acc = torch.empty((R,), dtype=torch.float32, device=dev)
    wts = torch.empty((R, S), dtype=torch.float32, device=dev)
    depth = torch.empty((R,), dtype=torch.float32, device=dev)
    pmax = torch.empty((R, 3), dtype=torch.float32, device=dev) if want_pts_max else None
    noise = _f32(noise) if noise is not None else None
    with torch.cuda.device(dev):
        check(_lib.load().nfb_composite_fwd(ptr(raw), ptr(z_vals), dptr, pitch, ptr(noise), R, S, int(bool(white_bkgd)),
                                            ptr(rgb), ptr(disp), ptr(acc), ptr(wts), ptr(depth), ptr(pmax), stream()),
              "nfb_composite_fwd")
    del keep
    return (rgb, disp, acc, wts, depth, pmax) if want_pts_max else (rgb, disp, acc, wts, depth)


def composite_bwd(raw, z_vals, rays_or_d, noise, white_bkgd, g_rgb, g_disp, g_acc, g_weights, g_depth):
    raw, z_vals = _f32(raw), _f32(z_vals)
    R, S = z_vals.shape
    keep, dptr, pitch, _ = _rays_d_view(rays_or_d)
    g_raw = torch.empty_like(raw)
    gs = [None if g is None else _f32(g) for g in (g_rgb, g_disp, g_acc, g_weights, g_depth)]
    noise = _f32(noise) if noise is not None else None
    with torch.cuda.device(raw.device):
        check(_lib.load().nfb_composite_bwd(ptr(raw), ptr(z_vals), dptr, pitch, ptr(noise), R, S, int(bool(white_bkgd)),
                                            ptr(gs[0]), ptr(gs[1]), ptr(gs[2]), ptr(gs[3]), ptr(gs[4]), ptr(g_raw),
                                            stream()), "nfb_composite_bwd")
    del keep
    return g_raw


class CompositeFn(torch.autograd.Function):
    """Differentiable raw2outputs: gradient flows to `raw` only (z_vals / rays are constants of the render,
    as in the reference where z_samples is detached, run_nerf.py:394)."""

    @staticmethod
    def forward(ctx, raw, z_vals, rays_or_d, noise, white_bkgd):
        out = composite_fwd(raw, z_vals, rays_or_d, noise, white_bkgd)
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(raw, z_vals, rays_or_d, noise if noise is not None else torch.empty(0, device=raw.device))
        ctx.has_noise = noise is not None
        ctx.white = bool(white_bkgd)
        return out

    @staticmethod
    def backward(ctx, g_rgb, g_disp, g_acc, g_weights, g_depth):
        raw, z_vals, rays_or_d, noise = ctx.saved_tensors
        if all(g is None for g in (g_rgb, g_disp, g_acc, g_weights, g_depth)):
            return None, None, None, None, None
        g_raw = composite_bwd(raw, z_vals, rays_or_d, noise if ctx.has_noise else None, ctx.white,
                              g_rgb, g_disp, g_acc, g_weights, g_depth)
        return g_raw, None, None, None, None


# ------------------------------------------------------------------------------------------------
# hierarchical sampling
# ------------------------------------------------------------------------------------------------
def sample_pdf(bins, weights, n_samples, u=None, return_inds=False):
    """run_nerf_helpers.py:200-243. u=None means det=True (linspace)."""
    _require_cuda(bins, weights, u)
    bins, weights = _f32(bins), _f32(weights)
    R, nb = bins.shape
    assert weights.shape == (R, nb - 1)
    if u is not None:
        u = _f32(u)
        assert u.shape == (R, n_samples)
    out = torch.empty((R, n_samples), dtype=torch.float32, device=bins.device)
    inds = torch.empty((R, n_samples), dtype=torch.int32, device=bins.device) if return_inds else None
    with torch.cuda.device(bins.device):
        check(_lib.load().nfb_sample_pdf(ptr(bins), ptr(weights), nb - 1, ptr(u), R, nb, n_samples, ptr(out), ptr(inds),
                                         stream()), "nfb_sample_pdf")
    return (out, inds) if return_inds else out


def hierarchical(z_coarse, weights, n_importance, u=None, rng: Optional[tuple] = None):
    """(z_fine [R,Sc+N] ascending, z_samples [R,N], z_std [R]) — run_nerf.py:392-396, :412.  rng = (seed, offset): u drawn in
    the kernel (Philox) instead of read from HBM."""
    _require_cuda(z_coarse, weights, u)
    z_coarse, weights = _f32(z_coarse), _f32(weights)
    R, Sc = z_coarse.shape
    assert weights.shape == (R, Sc)
    if u is not None:
        u = _f32(u)
        assert u.shape == (R, n_importance)
    dev = z_coarse.device
    z_fine = torch.empty((R, Sc + n_importance), dtype=torch.float32, device=dev)
    z_samples = torch.empty((R, n_importance), dtype=torch.float32, device=dev)
    z_std = torch.empty((R,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        if rng is not None and u is None:
            check(_lib.load().nfb_hierarchical_rng(ptr(z_coarse), ptr(weights), int(rng[0]), int(rng[1]), R, Sc, n_importance,
                                                   ptr(z_fine), ptr(z_samples), ptr(z_std), stream()), "nfb_hierarchical_rng")
        else:
            check(_lib.load().nfb_hierarchical(ptr(z_coarse), ptr(weights), ptr(u), R, Sc, n_importance, ptr(z_fine),
                                               ptr(z_samples), ptr(z_std), stream()), "nfb_hierarchical")
    return z_fine, z_samples, z_std


# ------------------------------------------------------------------------------------------------
# fused bf16 MLP
# ------------------------------------------------------------------------------------------------
class FusedMLP:
    """Owns an nfb_mlp_t handle (pre-swizzled bf16 weights) for one NeRF network."""

    def __init__(self, D=8, W=256, input_ch=63, input_ch_views=27, skip=4, device=None):
        self.device = torch.device(device if device is not None else "cuda")
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(_lib.load().nfb_mlp_create(C.byref(h), D, W, input_ch, input_ch_views, skip), "nfb_mlp_create")
        self._h = h
        self.n_params = int(_lib.load().nfb_mlp_param_count(h))

    def update(self, flat_params: torch.Tensor) -> None:
        flat_params = _f32(flat_params)
        with torch.cuda.device(self.device):
            check(_lib.load().nfb_mlp_update(self._h, ptr(flat_params), flat_params.numel(), stream()), "nfb_mlp_update")

    def _run(self, mode, pts, dirs, rays, z_vals, R, S, nsteps=10, want_dbg=False):
        raw = torch.empty((R, S, 4), dtype=torch.float32, device=self.device)
        dbg = torch.zeros((R * S, 256), dtype=torch.float32, device=self.device) if want_dbg else None
        lib = _lib.load()
        with torch.cuda.device(self.device):
            if want_dbg or nsteps != 10:
                check(lib.nfb_mlp_fwd_debug(self._h, mode, ptr(pts), ptr(dirs), ptr(rays), ptr(z_vals), R, S, ptr(raw),
                                            nsteps, ptr(dbg), stream()), "nfb_mlp_fwd_debug")
            else:
                check(lib.nfb_mlp_fwd(self._h, mode, ptr(pts), ptr(dirs), ptr(rays), ptr(z_vals), R, S, ptr(raw), stream()),
                      "nfb_mlp_fwd")
        return (raw, dbg) if want_dbg else raw

    def forward_points(self, pts: torch.Tensor, dirs: torch.Tensor, **kw):
        """pts [R,S,3], dirs [R,3] -> raw [R,S,4]."""
        pts, dirs = _f32(pts), _f32(dirs)
        R, S = pts.shape[0], pts.shape[1]
        return self._run(0, pts, dirs, None, None, R, S, **kw)

    def forward_rays(self, rays: torch.Tensor, z_vals: torch.Tensor, **kw):
        """rays [R,11], z_vals [R,S] -> raw [R,S,4] with pts = o + d*z formed in-kernel."""
        rays, z_vals = _f32(rays), _f32(z_vals)
        assert rays.shape[1] == 11
        R, S = z_vals.shape
        return self._run(1, None, None, rays, z_vals, R, S, **kw)

    def status(self) -> None:
        """Synchronises the device; raises if a pipeline barrier of this network's kernels timed out, and clears the flag."""
        check(_lib.load().nfb_mlp_status(self._h), "nfb_mlp_status")

    def poll(self) -> None:
        """The same check without synchronising (one read of pinned host memory): raises once a kernel that has already
        finished reported a time-out.  render_path's view sink, train_step and GraphedTrainStep call it at their hand-over
        points; every launch of the network polls too and refuses to run on top of invalid results."""
        check(_lib.load().nfb_mlp_poll(self._h), "nfb_mlp_poll")

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.load().nfb_mlp_destroy(self._h)
                self._h = None
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------
# fp32 layer-wise path
# ------------------------------------------------------------------------------------------------
def embed(x: torch.Tensor, L: int, out: torch.Tensor, col0: int, row_repeat: int = 1) -> None:
    """Writes [x, sin(2^l x), cos(2^l x)]_l into out[:, col0:col0+3+6L] (run_nerf_helpers.py:36-50)."""
    x = _f32(x)
    M = out.shape[0]
    assert out.is_contiguous() and out.dtype == torch.float32
    assert x.shape[0] * row_repeat == M
    with torch.cuda.device(out.device):
        check(_lib.load().nfb_embed(ptr(x), M, L, ptr(out), out.shape[1], col0, row_repeat, stream()), "nfb_embed")


class EmbedFn(torch.autograd.Function):
    """Positional encoding with the gradient to its input (run_nerf_helpers.py:36-50 is differentiable in the reference):
    d/dx [x, sin(2^l x), cos(2^l x)]_l = g_x + sum_l 2^l (cos_l * g_sin_l - sin_l * g_cos_l), with sin_l / cos_l read back
    from the saved output.  Only pose / ray optimisation built on the drop-in names needs it; no NeRFail path does, so
    the backward is a handful of torch ops rather than a kernel."""

    @staticmethod
    def forward(ctx, x, L, row_repeat):
        x = _f32(x)
        out = torch.empty((x.shape[0] * row_repeat, 3 + 6 * L), dtype=torch.float32, device=x.device)
        embed(x, L, out, 0, row_repeat)
        ctx.save_for_backward(out)
        ctx.L, ctx.row_repeat, ctx.n = L, row_repeat, x.shape[0]
        return out

    @staticmethod
    def backward(ctx, g):
        (out,) = ctx.saved_tensors
        L = ctx.L
        g = _f32(g)
        gx = g[:, 0:3].clone()
        sc = out[:, 3:].reshape(-1, L, 2, 3)
        gg = g[:, 3:].reshape(-1, L, 2, 3)
        freq = (2.0 ** torch.arange(L, device=g.device, dtype=torch.float32)).reshape(1, L, 1)
        gx += (freq * (sc[:, :, 1] * gg[:, :, 0] - sc[:, :, 0] * gg[:, :, 1])).sum(1)
        if ctx.row_repeat > 1:
            gx = gx.reshape(ctx.n, ctx.row_repeat, 3).sum(1)
        return gx, None, None


def _view2d(t: torch.Tensor):
    """(pointer, pitch) of a 2-D fp32 view whose rows are contiguous (column slices of a row-major buffer)."""
    assert t.dim() == 2 and t.dtype == torch.float32 and t.is_cuda
    assert t.shape[1] == 1 or t.stride(1) == 1, "rows must be contiguous"
    pitch = t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])
    return t.data_ptr(), pitch


_ws_cache = {}


def _workspace(nbytes: int, device) -> torch.Tensor:
    key = (device.type, device.index)
    cur = _ws_cache.get(key)
    if cur is None or cur.numel() * 4 < nbytes:
        cur = torch.empty((max(nbytes, 1 << 20) + 3) // 4, dtype=torch.float32, device=device)
        _ws_cache[key] = cur
    return cur


def linear_fwd(x, weight, bias, relu: bool, out: Optional[torch.Tensor] = None):
    M, K = x.shape
    N = weight.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=x.device)
    xp, ldx = _view2d(x)
    wp, ldw = _view2d(weight)
    yp, ldy = _view2d(out)
    with torch.cuda.device(x.device):
        check(_lib.load().nfb_linear_fwd(xp, ldx, wp, ldw, ptr(bias) if bias is not None else None, M, N, K, int(relu), yp, ldy,
                                         stream()), "nfb_linear_fwd")
    return out


def linear_bwd_data(dy, y, relu: bool, weight, out: Optional[torch.Tensor] = None, accumulate=False):
    M, N = dy.shape
    K = weight.shape[1]
    if out is None:
        out = torch.empty((M, K), dtype=torch.float32, device=dy.device)
    dyp, lddy = _view2d(dy)
    yp, ldy = _view2d(y) if relu else (None, 0)
    wp, ldw = _view2d(weight)
    dxp, lddx = _view2d(out)
    with torch.cuda.device(dy.device):
        check(_lib.load().nfb_linear_bwd_data(dyp, lddy, yp, ldy, int(relu), wp, ldw, M, N, K, dxp, lddx, int(accumulate),
                                              stream()), "nfb_linear_bwd_data")
    return out


def linear_bwd_weight(dy, y, relu: bool, x, want_bias=True):
    M, N = dy.shape
    K = x.shape[1]
    dev = dy.device
    dW = torch.empty((N, K), dtype=torch.float32, device=dev)
    db = torch.empty((N,), dtype=torch.float32, device=dev) if want_bias else None
    lib = _lib.load()
    nbytes = int(lib.nfb_linear_bwd_weight_workspace(M, N, K))
    ws = _workspace(nbytes, dev)
    dyp, lddy = _view2d(dy)
    yp, ldy = _view2d(y) if relu else (None, 0)
    xp, ldx = _view2d(x)
    with torch.cuda.device(dev):
        check(lib.nfb_linear_bwd_weight(dyp, lddy, yp, ldy, int(relu), xp, ldx, M, N, K, ptr(dW), K, ptr(db), ptr(ws),
                                        ws.numel() * 4, stream()), "nfb_linear_bwd_weight")
    return dW, db


class LinearFn(torch.autograd.Function):
    """y = act(x W^T + b) with the fp32 CUDA-core kernels (addmm + relu of run_nerf_helpers.py:104-118)."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        x2 = x if (x.dim() == 2 and x.stride(1) == 1) else x.contiguous()
        w2 = weight if weight.is_contiguous() else weight.contiguous()
        y = linear_fwd(x2, w2, bias, relu)
        ctx.relu = bool(relu)
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x2, w2, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        # rows must be contiguous and the row pitch real: an expanded gradient (e.g. from .sum().backward()) has pitch 0
        if not ((dy.stride(1) == 1 or dy.shape[1] == 1) and (dy.shape[0] <= 1 or dy.stride(0) >= dy.shape[1])):
            dy = dy.contiguous()
        dx = dW = db = None
        if ctx.needs_input_grad[0]:
            dx = linear_bwd_data(dy, y, ctx.relu, w)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dW, db = linear_bwd_weight(dy, y, ctx.relu, x, want_bias=ctx.has_bias)
        return dx, dW, db, None


# ------------------------------------------------------------------------------------------------
# GaussNet
# ------------------------------------------------------------------------------------------------
def knn8(query: torch.Tensor, cand: torch.Tensor):
    """(dist [Q,8] fp32, idx [Q,8] int32): exact 8-NN, create_index_and_dist.py:126-145."""
    _require_cuda(query, cand)
    query, cand = _f32(query).reshape(-1, 3), _f32(cand).reshape(-1, 3)
    Q, Cn = query.shape[0], cand.shape[0]
    dist = torch.empty((Q, 8), dtype=torch.float32, device=query.device)
    idx = torch.empty((Q, 8), dtype=torch.int32, device=query.device)
    with torch.cuda.device(query.device):
        check(_lib.load().nfb_knn8(ptr(query), Q, ptr(cand), Cn, ptr(dist), None, ptr(idx), stream()), "nfb_knn8")
    return dist, idx


class KnnGrid:
    """Exact 8-NN accelerator for a fixed candidate set (the P base views of a data set, create_index_and_dist.py:96-108):
    256^3 Morton grid over the candidates' bounding box, built once on the device (nfb_knn_grid_build); `query` returns
    exactly what `knn8` returns (same fp32 distance expression, ties to the lower index)."""

    def __init__(self, cand: torch.Tensor):
        _require_cuda(cand)
        lib = _lib.load()
        self.cand = _f32(cand).reshape(-1, 3)
        self.C = self.cand.shape[0]
        dev = self.cand.device
        finite = torch.isfinite(self.cand).all(dim=1, keepdim=True)
        big = torch.finfo(torch.float32).max
        lo = torch.where(finite, self.cand, torch.full_like(self.cand, big)).min(0).values.cpu().numpy().astype(np.float32)
        hi = torch.where(finite, self.cand, torch.full_like(self.cand, -big)).max(0).values.cpu().numpy().astype(np.float32)
        extent = float(np.max(hi - lo))
        if not np.isfinite(extent) or extent <= 0.0:
            extent = 1.0
            lo = np.zeros(3, np.float32) if not np.all(np.isfinite(lo)) else lo
        self.h = float(np.float32(extent * (1.0 + 1e-4) / 256.0))          # 256 cells cover the longest axis with slack
        self.lo = np.ascontiguousarray(lo, dtype=np.float32)
        self.margin_abs = float(4e-6 * max(float(np.abs(lo).max()), float(np.abs(hi).max()), 1e-30))
        cells = int(lib.nfb_knn_grid_cells())
        self.sorted = torch.empty((self.C, 4), dtype=torch.float32, device=dev)
        self.cell_start = torch.empty(cells + 1, dtype=torch.int32, device=dev)
        work = torch.empty(cells + 4096, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            check(lib.nfb_knn_grid_build(ptr(self.cand), self.C, self.lo.ctypes.data, self.h, ptr(self.sorted), ptr(self.cell_start),
                                         ptr(work), stream()), "nfb_knn_grid_build")

    def _run(self, q, dist, idx_f, idx_i, stats=None):
        with torch.cuda.device(q.device):
            check(_lib.load().nfb_knn8_grid(ptr(q), q.shape[0], ptr(self.sorted), ptr(self.cell_start), self.lo.ctypes.data, self.h,
                                            self.margin_abs, ptr(dist), ptr(idx_f), ptr(idx_i), ptr(stats), stream()), "nfb_knn8_grid")

    def query(self, query: torch.Tensor, stats: torch.Tensor | None = None):
        """(dist [Q,8] fp32, idx [Q,8] int32); stats: optional int64 device scalar accumulating distance evaluations."""
        _require_cuda(query)
        q = _f32(query).reshape(-1, 3)
        dist = torch.empty((q.shape[0], 8), dtype=torch.float32, device=q.device)
        idx = torch.empty((q.shape[0], 8), dtype=torch.int32, device=q.device)
        self._run(q, dist, None, idx, stats)
        return dist, idx

    def query_dist_idx(self, query_hw3: torch.Tensor) -> torch.Tensor:
        """float32 [2,H,W,8] = cat([dist, idx]), the reference's on-disk layout (create_index_and_dist.py:148-151)."""
        H, W = query_hw3.shape[0], query_hw3.shape[1]
        q = _f32(query_hw3).reshape(-1, 3)
        out = torch.empty((2, H * W, 8), dtype=torch.float32, device=q.device)
        self._run(q, out[0], out[1], None)
        return out.reshape(2, H, W, 8)


def knn8_dist_idx(query_hw3: torch.Tensor, cand: torch.Tensor) -> torch.Tensor:
    """The reference's on-disk layout: float32 [2,H,W,8] = cat([dist, idx]) (create_index_and_dist.py:148-151)."""
    H, W = query_hw3.shape[0], query_hw3.shape[1]
    q = _f32(query_hw3).reshape(-1, 3)
    cand = _f32(cand).reshape(-1, 3)
    out = torch.empty((2, H * W, 8), dtype=torch.float32, device=q.device)
    with torch.cuda.device(q.device):
        check(_lib.load().nfb_knn8(ptr(q), q.shape[0], ptr(cand), cand.shape[0], out[0].data_ptr(), out[1].data_ptr(), None,
                                   stream()), "nfb_knn8")
    return out.reshape(2, H, W, 8)


def gauss_weights(dist_idx: torch.Tensor, c: float) -> torch.Tensor:
    """[B,2,H,W,8] dist/idx -> [B,2,H,W,8] weight/idx (model/GaussNet.py:169-186)."""
    _require_cuda(dist_idx)
    di = _f32(dist_idx)
    B = di.shape[0]
    HW = di.shape[2] * di.shape[3]
    out = torch.empty_like(di)
    with torch.cuda.device(di.device):
        check(_lib.load().nfb_gauss_weights(ptr(di), B, HW, float(c), ptr(out), stream()), "nfb_gauss_weights")
    return out


def gauss_gather_fwd(table, w_idx, ori_u8, eps, minmax=None):
    T = table.numel() // 4
    B = w_idx.shape[0]
    HW = w_idx.shape[2] * w_idx.shape[3]
    shape = (B, w_idx.shape[2], w_idx.shape[3], 4)
    x = torch.empty(shape, dtype=torch.float32, device=table.device)
    x_rgba = torch.empty(shape, dtype=torch.float32, device=table.device)
    with torch.cuda.device(table.device):
        check(_lib.load().nfb_gauss_gather_fwd(ptr(table), T, ptr(w_idx), ptr(ori_u8), B, HW,
                                               -1.0 if eps is None else float(eps), ptr(x), ptr(x_rgba), ptr(minmax),
                                               stream()), "nfb_gauss_gather_fwd")
    return x, x_rgba


def gauss_scatter_bwd(g_x, g_xrgba, x, w_idx, ori_u8, eps, table_shape, out=None):
    """grad w.r.t. the [P,H,W,4] table; accumulates into `out` when given (e.g. across the views of an iteration)."""
    T = int(np.prod(table_shape)) // 4
    B = w_idx.shape[0]
    HW = w_idx.shape[2] * w_idx.shape[3]
    g_table = out if out is not None else torch.zeros(table_shape, dtype=torch.float32, device=w_idx.device)
    assert g_table.is_contiguous() and g_table.numel() == T * 4
    g_x = None if g_x is None else _f32(g_x)
    g_xrgba = None if g_xrgba is None else _f32(g_xrgba)
    with torch.cuda.device(w_idx.device):
        check(_lib.load().nfb_gauss_scatter_bwd(ptr(g_x), ptr(g_xrgba), ptr(x), ptr(w_idx), ptr(ori_u8), B, HW,
                                                -1.0 if eps is None else float(eps), T, ptr(g_table), stream()),
              "nfb_gauss_scatter_bwd")
    return g_table


class _GaussScatterFn(torch.autograd.Function):
    """g_table = J^T (g_x, g_xrgba).  Linear in (g_x, g_xrgba): its own backward is the gather with the same
    masks, which keeps gauss_net usable under create_graph=True (deepfool.py:76-77)."""

    @staticmethod
    def forward(ctx, g_x, g_xrgba, x, w_idx, ori_u8, eps, table_shape):
        ctx.save_for_backward(x, w_idx, ori_u8)
        ctx.eps = eps
        ctx.has = (g_x is not None, g_xrgba is not None)
        return gauss_scatter_bwd(g_x, g_xrgba, x, w_idx, ori_u8, eps, table_shape)

    @staticmethod
    def backward(ctx, gg_table):
        x, w_idx, ori_u8 = ctx.saved_tensors
        # d/d(g_x) <g_table, gg> = gather(gg) ; d/d(g_xrgba) = mask-scaled gather(gg)
        gx, _ = gauss_gather_fwd(_f32(gg_table).reshape(-1, 4), w_idx, ori_u8, None)
        g_gx = gx if ctx.has[0] else None
        g_gxrgba = None
        if ctx.has[1]:
            alpha = x[..., 3:4] / 255.0
            pr = x[..., :3] * alpha
            orif = ori_u8.reshape(x.shape).float()
            keep = orif[..., 3:4] > 0
            if ctx.eps is not None:
                keep = keep & (pr >= -ctx.eps) & (pr <= ctx.eps)
                pr = pr.clamp(-ctx.eps, ctx.eps)
            v = orif[..., :3] + pr
            keep = keep & (v >= 0) & (v <= 255)
            rgb = torch.where(keep, gx[..., :3] * alpha + gx[..., 3:4] * x[..., :3] / 255.0, torch.zeros_like(pr))
            g_gxrgba = torch.cat([rgb, torch.zeros_like(alpha)], -1)
        return g_gx, g_gxrgba, None, None, None, None, None


class GaussGatherFn(torch.autograd.Function):
    """(x, x_rgba) = gather/composite of model/GaussNet.py:53-119; backward = one red.global.add.v4.f32 per (pixel, neighbour)
    into the L2-resident table (the warp-aggregated variant measured slower on B200 and is opt-in, csrc/gauss.cu)."""

    @staticmethod
    def forward(ctx, spatial_rgb, w_idx, ori_u8, eps, minmax):
        table = _f32(spatial_rgb).reshape(-1, 4)
        x, x_rgba = gauss_gather_fwd(table, w_idx, ori_u8, eps, minmax)
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(x, w_idx, ori_u8)
        ctx.eps = eps
        ctx.table_shape = tuple(spatial_rgb.shape)
        return x, x_rgba

    @staticmethod
    def backward(ctx, g_x, g_xrgba):
        x, w_idx, ori_u8 = ctx.saved_tensors
        if g_x is None and g_xrgba is None:
            return None, None, None, None, None
        g = _GaussScatterFn.apply(g_x, g_xrgba, x, w_idx, ori_u8, ctx.eps, ctx.table_shape)
        return g, None, None, None, None


def philox_uniform(seed: int, offset: int, stream_id: int, n: int, device=None) -> torch.Tensor:
    """The n uniform numbers of Philox stream `stream_id` that the kernel-side draws consume (nfb_philox_uniform)."""
    out = torch.empty(n, dtype=torch.float32, device=torch.device(device if device is not None else "cuda"))
    with torch.cuda.device(out.device):
        check(_lib.load().nfb_philox_uniform(int(seed), int(offset), int(stream_id), n, ptr(out), stream()), "nfb_philox_uniform")
    return out


def next_philox(device=None) -> tuple:
    """(seed, offset) for one kernel-side draw, taken from torch's default CPU generator (two 62-bit integers: no GPU
    launch, no synchronisation), so torch.manual_seed makes a stochastic render reproducible exactly as it does for the
    torch.rand draws this replaces, and consecutive calls get unrelated streams."""
    k = torch.randint(0, 1 << 62, (2,), dtype=torch.int64)
    return int(k[0]), int(k[1])


def render_rays_fused(coarse: "FusedMLP", fine: Optional["FusedMLP"], rays: torch.Tensor, N_samples: int, N_importance: int,
                      lindisp: bool, white_bkgd: bool, t_rand=None, u=None, want_pts_max: bool = False,
                      rng: Optional[tuple] = None) -> dict:
    """nfb_render_rays_fwd: the whole no-grad kernel sequence of render_rays (run_nerf.py:308-418) for one ray batch in one
    C call (coarse depths, fused MLP, compositing, resampling + merge, fused MLP, compositing [+ pts_max])."""
    lib = _lib.load()
    rays = _f32(rays)
    R = rays.shape[0]
    dev = rays.device
    f = lambda *sh: torch.empty(sh, dtype=torch.float32, device=dev)
    out = {"rgb_map": f(R, 3), "disp_map": f(R), "acc_map": f(R)}
    two = N_importance > 0
    if two:
        out.update(rgb0=f(R, 3), disp0=f(R), acc0=f(R), z_std=f(R))
    if want_pts_max:
        out["pts_max"] = f(R, 3)
    if R == 0:
        return out
    nbytes = int(lib.nfb_render_rays_workspace_bytes(R, N_samples, N_importance))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    t_rand = _f32(t_rand) if t_rand is not None else None
    u = _f32(u) if u is not None else None
    with torch.cuda.device(dev):
        check(lib.nfb_render_rays_fwd(coarse._h, fine._h if fine is not None else None, ptr(rays), R, N_samples, N_importance,
                                      int(bool(lindisp)), int(bool(white_bkgd)), ptr(t_rand), ptr(u),
                                      int(rng is not None), int(rng[0]) if rng else 0, int(rng[1]) if rng else 0,
                                      ptr(out["rgb_map"]), ptr(out["disp_map"]), ptr(out["acc_map"]),
                                      ptr(out.get("rgb0")), ptr(out.get("disp0")), ptr(out.get("acc0")), ptr(out.get("z_std")),
                                      ptr(out.get("pts_max")), ptr(ws), nbytes, stream()), "nfb_render_rays_fwd")
    return out


class RenderRaysTrainFn(torch.autograd.Function):
    """render_rays under autograd as TWO C calls: nfb_render_rays_train_fwd (coarse depths, training forward of the coarse
    network, compositing, detached resampling, training forward of the fine network, compositing) and nfb_render_rays_bwd
    (compositing backward, data-gradient chain and grouped weight gradient of each network).  The same kernels in the same
    order as the per-op autograd path (NERFAIL_B200_RENDER_RAYS=ops), without ~50 Python-level launches per step.
    Gradients flow to the two networks' parameters only (rays and depths are constants of the step, run_nerf.py:394).
    Returns (rgb, disp, acc, rgb0, disp0, acc0, z_std, raw): z_std and raw are not differentiable."""

    @staticmethod
    def forward(ctx, net_c, net_f, rays, n_samples, n_importance, lindisp, white_bkgd, t_rand, u, *params):
        lib = _lib.load()
        fc, ff = net_c.fused(), net_f.fused()
        rays = _f32(rays)
        R = rays.shape[0]
        dev = rays.device
        Sf = n_samples + n_importance
        f = lambda *sh: torch.empty(sh, dtype=torch.float32, device=dev)
        rgb, disp, acc, rgb0, disp0, acc0, z_std = f(R, 3), f(R), f(R), f(R, 3), f(R), f(R), f(R)
        nbytes = int(lib.nfb_render_rays_train_workspace_bytes(R, n_samples, n_importance))
        ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=dev)
        t_rand = _f32(t_rand) if t_rand is not None else None
        u = _f32(u) if u is not None else None
        if R > 0:
            with torch.cuda.device(dev):
                check(lib.nfb_render_rays_train_fwd(fc._h, ff._h, ptr(rays), R, n_samples, n_importance, int(bool(lindisp)),
                                                    int(bool(white_bkgd)), ptr(t_rand), ptr(u), ptr(rgb), ptr(disp), ptr(acc),
                                                    ptr(rgb0), ptr(disp0), ptr(acc0), ptr(z_std), ptr(ws), nbytes, stream()),
                      "nfb_render_rays_train_fwd")
        off = int(lib.nfb_render_rays_train_raw_offset(R, n_samples, n_importance)) if R > 0 else 0
        raw = ws[off:off + R * Sf * 16].view(torch.float32).view(R, Sf, 4)
        ctx.nets, ctx.fused = (net_c, net_f), (fc, ff)
        ctx.cfg = (R, n_samples, n_importance, bool(white_bkgd), nbytes)
        ctx.n_c = len(net_c.ordered_params())
        ctx.shapes = [tuple(p.shape) for p in params]
        ctx.save_for_backward(rays, ws)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(z_std, raw)
        return rgb, disp, acc, rgb0, disp0, acc0, z_std, raw

    @staticmethod
    def backward(ctx, g_rgb, g_disp, g_acc, g_rgb0, g_disp0, g_acc0, _g_std, _g_raw):
        lib = _lib.load()
        rays, ws = ctx.saved_tensors
        R, n_samples, n_importance, white, nbytes = ctx.cfg
        dev = rays.device
        flats = []
        roots = [getattr(net, "_grad_sink_root", None) for net in ctx.nets]
        shared_root = roots[0] is not None and roots[0] is roots[1]
        if shared_root:                 # both networks' sinks are slices of one buffer (dist.PeerAdam): one fill
            roots[0].zero_()
        for net, fused in zip(ctx.nets, ctx.fused):
            sink = getattr(net, "_grad_sink", None)
            if sink is not None:        # data-parallel: the flat gradient lives in NVLink peer memory (dist.PeerAdam)
                if not shared_root:
                    sink.zero_()
                flats.append(sink)
            else:
                flats.append(torch.zeros(fused.n_params, dtype=torch.float32, device=dev))
        gs = [None if g is None else _f32(g) for g in (g_rgb, g_disp, g_acc, g_rgb0, g_disp0, g_acc0)]
        if R > 0 and any(g is not None for g in gs):
            with torch.cuda.device(dev):
                check(lib.nfb_render_rays_bwd(ctx.fused[0]._h, ctx.fused[1]._h, ptr(rays), R, n_samples, n_importance, int(white),
                                              ptr(gs[0]), ptr(gs[1]), ptr(gs[2]), ptr(gs[3]), ptr(gs[4]), ptr(gs[5]),
                                              ptr(flats[0]), ptr(flats[1]), ptr(ws), nbytes, stream()), "nfb_render_rays_bwd")
        grads, k = [], 0
        for i, shp in enumerate(ctx.shapes):        # views of the flat gradients in state_dict order: coarse first, then fine
            if i == ctx.n_c:
                k = 0
            flat = flats[0] if i < ctx.n_c else flats[1]
            n = math.prod(shp)
            grads.append(flat[k:k + n].view(shp))
            k += n
        return (None,) * 9 + tuple(grads)


class RgbaToChwFn(torch.autograd.Function):
    """Classifier input of model/GaussNet.py:121-145: [B,H,W,4] RGBA -> [B,3,H,W] RGB, `fill` where alpha (channel 3 of
    the image itself) is 0.  Backward = ChwToRgbaFn, whose backward is this op with fill 0: differentiable twice."""

    @staticmethod
    def forward(ctx, img, fill):
        img = _f32(img)
        B, H, W, _ = img.shape
        out = torch.empty((B, 3, H, W), dtype=torch.float32, device=img.device)
        with torch.cuda.device(img.device):
            check(_lib.load().nfb_rgba_to_chw(ptr(img), None, None, B, H * W, float(fill), ptr(out), stream()), "nfb_rgba_to_chw")
        ctx.save_for_backward(img)
        return out

    @staticmethod
    def backward(ctx, g_out):
        (img,) = ctx.saved_tensors
        return ChwToRgbaFn.apply(g_out, img), None


class ChwToRgbaFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, g_out, alpha_src):
        g_out = _f32(g_out)
        B, _, H, W = g_out.shape
        g_img = torch.empty((B, H, W, 4), dtype=torch.float32, device=g_out.device)
        with torch.cuda.device(g_out.device):
            check(_lib.load().nfb_chw_to_rgba(ptr(g_out), ptr(alpha_src), B, H * W, ptr(g_img), stream()), "nfb_chw_to_rgba")
        ctx.save_for_backward(alpha_src)
        return g_img

    @staticmethod
    def backward(ctx, gg_img):
        (alpha_src,) = ctx.saved_tensors
        gg = _f32(gg_img)
        B, H, W, _ = gg.shape
        out = torch.empty((B, 3, H, W), dtype=torch.float32, device=gg.device)
        with torch.cuda.device(gg.device):
            check(_lib.load().nfb_rgba_to_chw(ptr(gg), None, ptr(alpha_src), B, H * W, 0.0, ptr(out), stream()), "nfb_rgba_to_chw")
        return out, None


# ---- classifier input with the bilinear Resize fused behind the RGBA -> CHW conversion (GaussNet.py:121-154) ----
_resize_tables = {}


def resize_tables(in_size: int, out_size: int, antialias: bool, transposed: bool, device):
    """(start int32 [n], count int32 [n], weights fp32 [n, maxk], maxk) of one axis on `device`, cached: the separable
    bilinear resampling of torchvision Resize / ATen upsample_bilinear2d(_aa), computed on the host by nfb_resize_weights."""
    device = torch.device(device)
    key = (in_size, out_size, bool(antialias), bool(transposed), device.type, device.index)
    hit = _resize_tables.get(key)
    if hit is not None:
        return hit
    lib = _lib.load()
    maxk = int(lib.nfb_resize_max_taps(in_size, out_size, int(antialias), int(transposed)))
    n = in_size if transposed else out_size
    start, count = np.zeros(n, np.int32), np.zeros(n, np.int32)
    w = np.zeros((n, maxk), np.float32)
    check(lib.nfb_resize_weights(in_size, out_size, int(antialias), int(transposed), maxk, start.ctypes.data, count.ctypes.data,
                                 w.ctypes.data), "nfb_resize_weights")
    hit = (torch.from_numpy(start).to(device), torch.from_numpy(count).to(device), torch.from_numpy(w).to(device), maxk)
    _resize_tables[key] = hit
    return hit


def default_resize_antialias() -> bool:
    """What `torchvision.transforms.Resize([s, s])` does to a float tensor in the torchvision that is installed: antialias
    defaults to True from 0.17 on; the 0.15 the reference pins (README.md:56) warns and does NOT antialias tensors.
    NERFAIL_B200_RESIZE_ANTIALIAS=0/1 overrides."""
    env = os.environ.get("NERFAIL_B200_RESIZE_ANTIALIAS")
    if env is not None:
        return env == "1"
    try:
        import torchvision
        major, minor = (int(v) for v in torchvision.__version__.split(".")[:2])
        return (major, minor) >= (0, 17)
    except Exception:
        return True


def _rgba_to_chw_resized(img_f32, img_u8, alpha_src, alpha_batch, B, H, W, size, fill, antialias):
    src = img_f32 if img_f32 is not None else img_u8
    dev = src.device
    ys, yc, yw, yk = resize_tables(H, size, antialias, False, dev)
    xs, xc, xw, xk = resize_tables(W, size, antialias, False, dev)
    out = torch.empty((B, 3, size, size), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load().nfb_rgba_to_chw_resized(ptr(img_f32), ptr(img_u8), ptr(alpha_src), alpha_batch, B, H, W, size, size,
                                                  float(fill), ptr(ys), ptr(yc), ptr(yw), yk, ptr(xs), ptr(xc), ptr(xw), xk,
                                                  ptr(out), stream()), "nfb_rgba_to_chw_resized")
    return out


def chw_resized_to_rgba(g_out: torch.Tensor, alpha_src: torch.Tensor, H: int, W: int, antialias: bool) -> torch.Tensor:
    """Adjoint of the fused conversion + resize: g_out [NB,3,S,S] -> g_img [NB,H,W,4]; alpha_src [B,H,W,4] with NB a multiple
    of B (cotangent n uses image n % B — NC class gradients of the same B images in one launch)."""
    g_out = _f32(g_out)
    NB, _, S, _ = g_out.shape
    Bm = alpha_src.shape[0]
    assert NB % Bm == 0
    dev = g_out.device
    ys, yc, yw, yk = resize_tables(H, S, antialias, True, dev)
    xs, xc, xw, xk = resize_tables(W, S, antialias, True, dev)
    g_img = torch.empty((NB, H, W, 4), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load().nfb_chw_resized_to_rgba(ptr(g_out), ptr(alpha_src), Bm, NB, H, W, S, S, ptr(ys), ptr(yc), ptr(yw), yk,
                                                  ptr(xs), ptr(xc), ptr(xw), xk, ptr(g_img), stream()), "nfb_chw_resized_to_rgba")
    return g_img


class RgbaToChwResizedFn(torch.autograd.Function):
    """[B,H,W,4] RGBA -> [B,3,S,S]: white where alpha == 0, NCHW, bilinear Resize([S,S]) — one kernel (GaussNet.py:121-154).
    Backward = ChwResizedToRgbaFn (the adjoint gather), whose backward is this op with fill 0: differentiable twice."""

    @staticmethod
    def forward(ctx, img, fill, size, antialias):
        img = _f32(img)
        B, H, W, _ = img.shape
        ctx.save_for_backward(img)
        ctx.antialias = antialias
        return _rgba_to_chw_resized(img, None, None, 1, B, H, W, size, fill, antialias)

    @staticmethod
    def backward(ctx, g_out):
        (img,) = ctx.saved_tensors
        return ChwResizedToRgbaFn.apply(g_out, img, ctx.antialias), None, None, None


class ChwResizedToRgbaFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, g_out, alpha_src, antialias):
        ctx.save_for_backward(alpha_src)
        ctx.antialias, ctx.size = antialias, g_out.shape[-1]
        return chw_resized_to_rgba(g_out, alpha_src, alpha_src.shape[1], alpha_src.shape[2], antialias)

    @staticmethod
    def backward(ctx, gg_img):
        (alpha_src,) = ctx.saved_tensors
        gg = _f32(gg_img)
        B, H, W, _ = gg.shape
        return _rgba_to_chw_resized(gg, None, alpha_src, alpha_src.shape[0], B, H, W, ctx.size, 0.0, ctx.antialias), None, None


def rgba_u8_to_chw_resized(img_u8: torch.Tensor, size: int, antialias: bool, fill: float = 255.0) -> torch.Tensor:
    """The same for the original uint8 image (no gradient)."""
    img_u8 = img_u8.contiguous()
    B, H, W, _ = img_u8.shape
    return _rgba_to_chw_resized(None, img_u8, None, 1, B, H, W, size, fill, antialias)


def gauss_scatter_bwd_batched(g_xrgba, x, w_idx, ori_u8, eps, table_shape):
    """[NC,B,H,W,4] cotangents of x_rgba -> [NC,*table_shape] gradients w.r.t. the perturbation table in one launch
    (nfb_gauss_scatter_bwd_batched: DeepFool's per-class gradients, deepfool.py:72-86)."""
    g_xrgba = _f32(g_xrgba)
    NC = g_xrgba.shape[0]
    T = int(np.prod(table_shape)) // 4
    B = w_idx.shape[0]
    HW = w_idx.shape[2] * w_idx.shape[3]
    assert g_xrgba.numel() == NC * B * HW * 4
    g_table = torch.zeros((NC,) + tuple(table_shape), dtype=torch.float32, device=w_idx.device)
    with torch.cuda.device(w_idx.device):
        check(_lib.load().nfb_gauss_scatter_bwd_batched(ptr(g_xrgba), NC, ptr(x), ptr(w_idx), ptr(ori_u8), B, HW,
                                                        -1.0 if eps is None else float(eps), T, ptr(g_table), stream()),
              "nfb_gauss_scatter_bwd_batched")
    return g_table


def rgba_u8_to_chw(img_u8: torch.Tensor, fill: float = 255.0) -> torch.Tensor:
    """The same conversion for the original uint8 image (no gradient): [B,H,W,4] uint8 -> [B,3,H,W] float32."""
    img_u8 = img_u8.contiguous()
    B, H, W, _ = img_u8.shape
    out = torch.empty((B, 3, H, W), dtype=torch.float32, device=img_u8.device)
    with torch.cuda.device(img_u8.device):
        check(_lib.load().nfb_rgba_to_chw(None, ptr(img_u8), None, B, H * W, float(fill), ptr(out), stream()), "nfb_rgba_to_chw")
    return out


# ------------------------------------------------------------------------------------------------
# bf16 tile images + tensor-core weight gradient
# ------------------------------------------------------------------------------------------------
def to_tile_image(x: torch.Tensor) -> torch.Tensor:
    """[M, C] (M % 128 == 0, C % 64 == 0) -> bf16 tile image [M/128, C/64, 128, 64] with the 128-byte swizzle
    (16-byte unit u of row r is stored at unit u ^ (r % 8)); test / interchange helper, pure indexing."""
    M, Cc = x.shape
    assert M % 128 == 0 and Cc % 64 == 0
    t = x.to(torch.bfloat16).reshape(M // 128, 128, Cc // 64, 8, 8).permute(0, 2, 1, 3, 4)      # tile, chunk, row, unit, elem
    r = torch.arange(128, device=x.device).reshape(1, 1, 128, 1, 1)
    u = torch.arange(8, device=x.device).reshape(1, 1, 1, 8, 1)
    src_unit = (u ^ (r & 7)).expand(t.shape[0], t.shape[1], 128, 8, 8)
    out = torch.gather(t, 3, src_unit)          # out[.., r, p, :] = t[.., r, p ^ (r&7), :]  (XOR is an involution)
    return out.reshape(M // 128, Cc // 64, 128, 64).contiguous()


def wgrad_bf16(dy_img: torch.Tensor, x_img: torch.Tensor, want_bias=True):
    """dy_img [T, ndy, 128, 64] bf16, x_img [T, nx, 128, 64] bf16 (tile images; may be views with a tile pitch) ->
    (dW [64*ndy, 64*nx] fp32, db [64*ndy] fp32 or None).  One product of nfb_mlp_bwd_weights' grouped launch."""
    lib = _lib.load()
    T, ndy = dy_img.shape[0], dy_img.shape[1]
    nx = x_img.shape[1]
    assert dy_img.dtype == torch.bfloat16 and x_img.dtype == torch.bfloat16 and x_img.shape[0] == T
    assert dy_img.stride(1) == 8192 and x_img.stride(1) == 8192 and dy_img.stride(3) == 1 and x_img.stride(3) == 1
    dev = dy_img.device
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    dW = torch.zeros((64 * ndy, 64 * nx), dtype=torch.float32, device=dev)
    db = torch.zeros((64 * ndy,), dtype=torch.float32, device=dev) if want_bias else None
    with torch.cuda.device(dev):
        check(lib.nfb_wgrad_bf16(dy_img.data_ptr(), dy_img.stride(0) * 2, ndy, x_img.data_ptr(), x_img.stride(0) * 2, nx, T,
                                 ptr(dW), 64 * nx, 0, 64 * nx, 0, 64 * ndy, ptr(db), ptr(status), stream()), "nfb_wgrad_bf16")
    if int(status.item()) != 0:
        raise RuntimeError("nfb_wgrad_bf16: pipeline barrier timed out")
    return dW, db


def from_tile_image(img: torch.Tensor) -> torch.Tensor:
    """Inverse of to_tile_image: [T, C, 128, 64] swizzled bf16 -> [T*128, C*64] float32 (test helper)."""
    T, Cn = img.shape[0], img.shape[1]
    t = img.reshape(T, Cn, 128, 8, 8)
    r = torch.arange(128, device=img.device).reshape(1, 1, 128, 1, 1)
    u = torch.arange(8, device=img.device).reshape(1, 1, 1, 8, 1)
    un = torch.gather(t, 3, (u ^ (r & 7)).expand(T, Cn, 128, 8, 8))
    return un.permute(0, 2, 1, 3, 4).reshape(T * 128, Cn * 64).float()


class FusedMLPTrainFn(torch.autograd.Function):
    """raw = NeRF(rays, z) on tensor cores with autograd: forward = nfb_mlp_fwd_train (saves bf16 activations as tile
    images), backward = nfb_mlp_bwd_data (dY images) + nfb_mlp_bwd_weights (all 16 dW products in one grouped launch).
    Gradients flow to the network parameters only (rays / depths are constants of the step, run_nerf.py:394)."""

    @staticmethod
    def forward(ctx, net, rays, z_vals, *params):
        lib = _lib.load()
        fused = net.fused()
        rays, z_vals = _f32(rays), _f32(z_vals)
        R, S = z_vals.shape
        M = R * S
        T = int(lib.nfb_mlp_train_tiles(M))
        dev = rays.device
        act = torch.empty((T, 40, 128, 64), dtype=torch.bfloat16, device=dev)
        mask = torch.empty((T, 9, 8, 128), dtype=torch.int32, device=dev)
        raw = torch.empty((R, S, 4), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.nfb_mlp_fwd_train(fused._h, ptr(rays), ptr(z_vals), R, S, ptr(raw), ptr(act), ptr(mask), stream()),
                  "nfb_mlp_fwd_train")
        ctx.fused, ctx.M, ctx.T, ctx.net = fused, M, T, net
        ctx.shapes = [tuple(p.shape) for p in params]
        ctx.n_params = sum(math.prod(sh) for sh in ctx.shapes)
        ctx.save_for_backward(act, mask)
        return raw

    @staticmethod
    def backward(ctx, g_raw):
        lib = _lib.load()
        act, mask = ctx.saved_tensors
        M, T = ctx.M, ctx.T
        dev = act.device
        g_raw = _f32(g_raw).reshape(M, 4)
        dy = torch.empty((T, 39, 128, 64), dtype=torch.bfloat16, device=dev)
        sink = getattr(ctx.net, "_grad_sink", None)
        if sink is not None:        # data-parallel: the flat gradient lives in NVLink peer memory (dist.PeerAdam)
            grad = sink
            grad.zero_()
        else:
            grad = torch.zeros(ctx.n_params, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            if os.environ.get("NERFAIL_B200_BWD", "serial") != "overlap":
                check(lib.nfb_mlp_bwd_data(ctx.fused._h, ptr(g_raw), M, ptr(mask), ptr(dy), stream()), "nfb_mlp_bwd_data")
                check(lib.nfb_mlp_bwd_weights(ctx.fused._h, ptr(act), ptr(dy), ptr(g_raw), M, ptr(grad), stream()), "nfb_mlp_bwd_weights")
            else:
                # data-gradient and weight-gradient kernels side by side, dY handed over through L2 (nfb_mlp_bwd);
                # opt-in: measured slower than the serial pair on B200 (profiles/r01_train_bf16.md, "overlapped backward")
                ready = torch.empty(T, dtype=torch.int32, device=dev)
                check(lib.nfb_mlp_bwd(ctx.fused._h, ptr(g_raw), M, ptr(mask), ptr(act), ptr(dy), ptr(grad), ptr(ready),
                                      stream()), "nfb_mlp_bwd")
        # views of the flat gradient in state_dict order (the order FusedMLPTrainFn.apply received the parameters in)
        grads, o = [], 0
        for shp in ctx.shapes:
            n = math.prod(shp)
            grads.append(grad[o:o + n].view(shp))
            o += n
        return (None, None, None, *grads)
