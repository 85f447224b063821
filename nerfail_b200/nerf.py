"""Host-side mirror of nerf-pytorch's run_nerf_helpers.py for the B200 kernels.

Same names, arguments and state_dict keys as the reference
(Create_spatial_point_set/nerf_pytorch/run_nerf_helpers.py); the arithmetic runs in libnerfail_b200.so.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.nn as nn

from . import ops

# Misc (run_nerf_helpers.py:9-11) — host-side scalars, not on the kernel path
img2mse = lambda x, y: torch.mean((x - y) ** 2)
mse2psnr = lambda x: -10.0 * torch.log(x) / torch.log(torch.tensor([10.0], device=x.device))
to8b = lambda x: (255 * np.clip(x, 0, 1)).astype(np.uint8)


class Embedder:
    """Positional encoding (run_nerf_helpers.py:15-50): [x, sin(2^0 x), cos(2^0 x), ..., cos(2^(L-1) x)].

    Only the configuration the reference ever builds (get_embedder, :55-62: include_input, log sampling,
    sin/cos) is supported, which is what lets the encoding fuse into the first MLP layer.
    """

    def __init__(self, multires: int, input_dims: int = 3):
        if input_dims != 3:
            raise ValueError("nerfail_b200 Embedder supports 3-D inputs only")
        self.multires = int(multires)
        self.input_dims = 3
        self.out_dim = 3 + 6 * self.multires

    def embed(self, inputs: torch.Tensor, out: torch.Tensor | None = None, col0: int = 0, row_repeat: int = 1):
        """Writes the encoding of inputs [..., 3] into out[:, col0:col0+out_dim] (a fresh [M, out_dim] tensor when out is
        None).  Writing into a caller's buffer bypasses autograd, so it is refused for inputs that require grad: use
        __call__ / embed_grad, which are differentiable w.r.t. the inputs like the reference's torch Embedder."""
        x = inputs.reshape(-1, 3)
        if x.requires_grad and torch.is_grad_enabled():
            if out is not None:
                raise RuntimeError("nerfail_b200.Embedder.embed(out=...) is not differentiable w.r.t. its inputs; "
                                   "call the embedder (or embed_grad) to keep the gradient to points / rays")
            return ops.EmbedFn.apply(x, self.multires, row_repeat)
        if out is None:
            out = torch.empty((x.shape[0] * row_repeat, self.out_dim), dtype=torch.float32, device=x.device)
            col0 = 0
        ops.embed(x, self.multires, out, col0, row_repeat)
        return out

    def embed_grad(self, inputs: torch.Tensor, row_repeat: int = 1):
        """[M * row_repeat, out_dim] with autograd to the inputs (row i of the input fills rows i*row_repeat ...)."""
        return ops.EmbedFn.apply(inputs.reshape(-1, 3), self.multires, row_repeat)

    def __call__(self, inputs: torch.Tensor):
        lead = inputs.shape[:-1]
        return self.embed(inputs).reshape(*lead, self.out_dim)


def get_embedder(multires, i=0, device=None):
    """run_nerf_helpers.py:52-67. Returns (embed_fn, out_dim)."""
    if i == -1:
        return nn.Identity(), 3
    e = Embedder(multires)
    return e, e.out_dim


class NeRF(nn.Module):
    """run_nerf_helpers.py:71-123 with identical parameter names/shapes (checkpoints load unchanged).

    forward(x) takes already-embedded features [M, input_ch + input_ch_views] like the reference and runs the
    fp32 layer-wise kernels (differentiable).  The fused bf16 tcgen05 path is reached through
    run_network()/render_rays(), which hand raw points to `fused()` instead of embedding them first.
    """

    def __init__(self, D=8, W=256, input_ch=3, input_ch_views=3, output_ch=4, skips=[4], use_viewdirs=False):
        super().__init__()
        self.D, self.W = D, W
        self.input_ch, self.input_ch_views = input_ch, input_ch_views
        self.skips = list(skips)
        self.use_viewdirs = use_viewdirs
        self.pts_linears = nn.ModuleList(
            [nn.Linear(input_ch, W)]
            + [nn.Linear(W, W) if i not in self.skips else nn.Linear(W + input_ch, W) for i in range(D - 1)])
        self.views_linears = nn.ModuleList([nn.Linear(input_ch_views + W, W // 2)])
        if use_viewdirs:
            self.feature_linear = nn.Linear(W, W)
            self.alpha_linear = nn.Linear(W, 1)
            self.rgb_linear = nn.Linear(W // 2, 3)
        else:
            self.output_linear = nn.Linear(W, output_ch)
        self._fused = None
        self._fused_version = None

    # ---- fp32 layer-wise path (training / exact parity) ----
    def forward(self, x):
        x = x.reshape(-1, x.shape[-1])
        if not x.is_cuda:
            raise RuntimeError("nerfail_b200.NeRF runs on CUDA only")
        x = x.float()
        if x.stride(-1) != 1:
            x = x.contiguous()
        input_pts = x[:, : self.input_ch]
        input_views = x[:, self.input_ch: self.input_ch + self.input_ch_views]
        h = input_pts
        for i, lin in enumerate(self.pts_linears):
            h = ops.LinearFn.apply(h, lin.weight, lin.bias, True)
            if i in self.skips:
                h = torch.cat([input_pts, h], -1)
        if self.use_viewdirs:
            alpha = ops.LinearFn.apply(h, self.alpha_linear.weight, self.alpha_linear.bias, False)
            feature = ops.LinearFn.apply(h, self.feature_linear.weight, self.feature_linear.bias, False)
            h = torch.cat([feature, input_views], -1)
            for lin in self.views_linears:
                h = ops.LinearFn.apply(h, lin.weight, lin.bias, True)
            rgb = ops.LinearFn.apply(h, self.rgb_linear.weight, self.rgb_linear.bias, False)
            return torch.cat([rgb, alpha], -1)
        return ops.LinearFn.apply(h, self.output_linear.weight, self.output_linear.bias, False)

    # ---- fused bf16 tcgen05 path ----
    def fused_supported(self) -> bool:
        return (self.use_viewdirs and self.D == 8 and self.W == 256 and self.input_ch == 63
                and self.input_ch_views == 27 and self.skips == [4])

    def flat_params(self) -> torch.Tensor:
        """state_dict order expected by nfb_mlp_update."""
        parts = []
        for lin in self.pts_linears:
            parts += [lin.weight, lin.bias]
        parts += [self.views_linears[0].weight, self.views_linears[0].bias,
                  self.feature_linear.weight, self.feature_linear.bias,
                  self.alpha_linear.weight, self.alpha_linear.bias,
                  self.rgb_linear.weight, self.rgb_linear.bias]
        first = parts[0]
        if first.dtype == torch.float32 and first.is_contiguous():
            # parameters that already lie back to back in one buffer (dist.PeerAdam re-points them into peer memory in
            # exactly this order) are handed to the pack kernel as they are: no concatenation copy
            base, off, ok = first.untyped_storage().data_ptr(), first.storage_offset(), True
            for p in parts:
                if (p.untyped_storage().data_ptr() != base or p.storage_offset() != off or not p.is_contiguous()
                        or p.dtype != torch.float32):
                    ok = False
                    break
                off += p.numel()
            if ok:
                return torch.as_strided(first.detach(), (off - first.storage_offset(),), (1,), first.storage_offset())
        return torch.cat([p.detach().reshape(-1).float() for p in parts])

    def ordered_params(self):
        """Parameters in the state_dict / nfb_mlp_update order."""
        ps = []
        for lin in self.pts_linears:
            ps += [lin.weight, lin.bias]
        ps += [self.views_linears[0].weight, self.views_linears[0].bias, self.feature_linear.weight, self.feature_linear.bias,
               self.alpha_linear.weight, self.alpha_linear.bias, self.rgb_linear.weight, self.rgb_linear.bias]
        return ps

    def forward_rays_train(self, rays, z_vals):
        """raw [R,S,4] with autograd through the bf16 tensor-core training kernels (rays [R,11], z_vals [R,S])."""
        return ops.FusedMLPTrainFn.apply(self, rays, z_vals, *self.ordered_params())

    def invalidate_fused(self) -> None:
        """Forces the next fused() call to re-pack the bf16 weight image.  fused() notices optimizer steps and
        load_state_dict through the parameters' version counters; writes that bypass them — `p.data.copy_()`,
        `dist.broadcast(p.data, 0)`, EMA through `.data`, a raw-pointer kernel — do not bump the counter, so call this
        after any such write (nerfail_b200.dist.broadcast_params_ and GraphedTrainStep do)."""
        self._fused_version = None

    def close(self) -> None:
        """Releases the packed handle (device memory + its constant-memory entry) now instead of at garbage collection."""
        self._fused = None
        self._fused_version = None

    def _param_version(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def fused(self) -> "ops.FusedMLP":
        """The packed bf16 handle, re-packed whenever a parameter changed (optimizer.step, load_state_dict)."""
        if not self.fused_supported():
            raise RuntimeError("fused kernel supports D=8, W=256, multires 10/4, skips=[4], use_viewdirs=True only")
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("nerfail_b200.NeRF runs on CUDA only")
        ver = self._param_version()
        if self._fused is None or self._fused.device != dev:
            self._fused = ops.FusedMLP(device=dev)
            self._fused_version = None
        if self._fused_version != ver:
            self._fused.update(self.flat_params())
            self._fused_version = ver
        return self._fused


def train_precision() -> str:
    """Precision of the MLP under autograd: 'bf16' (fused tensor-core forward / data-gradient / weight-gradient kernels,
    mixed-precision gradients: 0.1-2.5 % relative) or 'fp32' (layer-wise CUDA-core kernels, gradients within 1e-3 of the
    reference, ~50x slower).  NERFAIL_B200_TRAIN selects it; unset, it follows NERFAIL_B200_MLP (default bf16), so one
    switch puts rendering and training on the exact-parity path."""
    v = os.environ.get("NERFAIL_B200_TRAIN")
    return v.lower() if v else mlp_precision()


def mlp_precision() -> str:
    """'bf16' (fused tcgen05, default for no-grad rendering) or 'fp32' (layer-wise kernels)."""
    return os.environ.get("NERFAIL_B200_MLP", "bf16").lower()


# Ray helpers -------------------------------------------------------------------------------------
def get_rays(H, W, K, c2w):
    """run_nerf_helpers.py:157-166. Returns (rays_o, rays_d) as [H,W,3] CUDA tensors."""
    dev = c2w.device if isinstance(c2w, torch.Tensor) and c2w.is_cuda else torch.device("cuda")
    rays = ops.get_ray_batch(H, W, K, c2w, 0.0, 1.0, device=dev)
    return rays[:, 0:3].reshape(H, W, 3), rays[:, 3:6].reshape(H, W, 3)


def get_rays_np(H, W, K, c2w):
    """run_nerf_helpers.py:169-176 (host-side numpy helper used by the training loop's batching)."""
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing="xy")
    dirs = np.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -np.ones_like(i)], -1)
    rays_d = np.sum(dirs[..., np.newaxis, :] * c2w[:3, :3], -1)
    rays_o = np.broadcast_to(c2w[:3, -1], np.shape(rays_d))
    return rays_o, rays_d


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    """run_nerf_helpers.py:179-197 — forward-facing scenes only; no NeRFail config uses it (all are blender,
    run_nerf.py:250-253), so it stays a handful of torch ops rather than a kernel."""
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t[..., None] * rays_d
    o0 = -1.0 / (W / (2.0 * focal)) * rays_o[..., 0] / rays_o[..., 2]
    o1 = -1.0 / (H / (2.0 * focal)) * rays_o[..., 1] / rays_o[..., 2]
    o2 = 1.0 + 2.0 * near / rays_o[..., 2]
    d0 = -1.0 / (W / (2.0 * focal)) * (rays_d[..., 0] / rays_d[..., 2] - rays_o[..., 0] / rays_o[..., 2])
    d1 = -1.0 / (H / (2.0 * focal)) * (rays_d[..., 1] / rays_d[..., 2] - rays_o[..., 1] / rays_o[..., 2])
    d2 = -2.0 * near / rays_o[..., 2]
    return torch.stack([o0, o1, o2], -1), torch.stack([d0, d1, d2], -1)


# Hierarchical sampling ---------------------------------------------------------------------------
def sample_pdf(bins, weights, N_samples, det=False, pytest=False):
    """run_nerf_helpers.py:200-243. Same signature; u comes from torch.rand (or numpy when pytest=True,
    exactly as the reference's :215-223 hook) and the CDF / search / lerp run in one kernel."""
    lead = bins.shape[:-1]
    b2 = bins.reshape(-1, bins.shape[-1])
    w2 = weights.reshape(-1, weights.shape[-1])
    u = None
    if pytest:
        np.random.seed(0)
        if not det:
            u = torch.tensor(np.random.rand(*(list(lead) + [N_samples])), dtype=torch.float32, device=bins.device)
    elif not det:
        u = torch.rand(list(lead) + [N_samples], device=bins.device)
    if u is not None:
        u = u.reshape(-1, N_samples)
    out = ops.sample_pdf(b2, w2, N_samples, u)
    return out.reshape(*lead, N_samples)
