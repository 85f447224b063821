"""Oracle restatement of the NeRF render path (TEST INFRASTRUCTURE — see oracle/__init__.py).

All citations are relative to the reference tree:
  H  = Create_spatial_point_set/nerf_pytorch/run_nerf_helpers.py
  R  = Create_spatial_point_set/nerf_pytorch/run_nerf.py
  C  = Create_spatial_point_set/nerf_to_coord.py
Everything is fp32 torch on the CPU, written functionally (weights come in as a dict of tensors).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ---------------------------------------------------------------------------------------------
# H:15-67  positional encoding
# ---------------------------------------------------------------------------------------------
def positional_encoding(x: Tensor, n_freqs: int) -> Tensor:
    """[x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)], each block as wide as x (H:24-50);
    the bands are 2**linspace(0, L-1, L) (H:32)."""
    bands = 2.0 ** torch.linspace(0.0, n_freqs - 1, steps=n_freqs)
    cols = [x]
    for b in bands:
        cols.append(torch.sin(x * b))
        cols.append(torch.cos(x * b))
    return torch.cat(cols, dim=-1)


# ---------------------------------------------------------------------------------------------
# H:71-123  the MLP, functional
# ---------------------------------------------------------------------------------------------
def nerf_mlp(sd: Dict[str, Tensor], feats: Tensor, input_ch: int = 63, input_ch_views: int = 27,
             depth: int = 8, skips=(4,)) -> Tensor:
    """NeRF.forward with use_viewdirs=True (H:100-121): eight relu layers with the encoded point re-concatenated
    IN FRONT of h after layer index 4 (H:106-107), sigma head and feature head without activation (H:110-111),
    one 283->128 relu view layer (H:114-116), rgb head (H:118); output [rgb(3), sigma(1)] (H:119)."""
    pts, views = feats[..., :input_ch], feats[..., input_ch:input_ch + input_ch_views]
    h = pts
    for i in range(depth):
        h = F.relu(F.linear(h, sd[f"pts_linears.{i}.weight"], sd[f"pts_linears.{i}.bias"]))
        if i in skips:
            h = torch.cat([pts, h], dim=-1)
    sigma = F.linear(h, sd["alpha_linear.weight"], sd["alpha_linear.bias"])
    feat = F.linear(h, sd["feature_linear.weight"], sd["feature_linear.bias"])
    hv = F.relu(F.linear(torch.cat([feat, views], dim=-1), sd["views_linears.0.weight"], sd["views_linears.0.bias"]))
    rgb = F.linear(hv, sd["rgb_linear.weight"], sd["rgb_linear.bias"])
    return torch.cat([rgb, sigma], dim=-1)


def query_network(sd, pts: Tensor, viewdirs: Tensor, l_pts: int = 10, l_dir: int = 4) -> Tensor:
    """run_network (R:37-51): encode the flattened points, broadcast each ray's direction to all of its
    samples (R:44-45), encode, concatenate, apply the MLP, restore [R,S,4]."""
    n_rays, n_samp = pts.shape[0], pts.shape[1]
    e_pts = positional_encoding(pts.reshape(-1, 3), l_pts)
    e_dir = positional_encoding(viewdirs[:, None, :].expand(n_rays, n_samp, 3).reshape(-1, 3), l_dir)
    out = nerf_mlp(sd, torch.cat([e_pts, e_dir], dim=-1), 3 + 6 * l_pts, 3 + 6 * l_dir)
    return out.reshape(n_rays, n_samp, 4)


# ---------------------------------------------------------------------------------------------
# H:157-166 + R:102-123  rays
# ---------------------------------------------------------------------------------------------
def camera_rays(H: int, W: int, K, c2w: Tensor, near: float, far: float) -> Tensor:
    """[H*W, 11] = origin, direction, near, far, unit view direction.
    dirs = ((i-cx)/fx, -(j-cy)/fy, -1) with i the column and j the row (H:158-161); rotated by c2w[:3,:3] as a
    multiply-then-sum over the last axis (H:163); origin = c2w[:3,3] (H:165); viewdirs = d/||d|| (R:108)."""
    cols = torch.arange(W, dtype=torch.float32)[None, :].expand(H, W)
    rows = torch.arange(H, dtype=torch.float32)[:, None].expand(H, W)
    fx, fy, cx, cy = float(K[0][0]), float(K[1][1]), float(K[0][2]), float(K[1][2])
    d_cam = torch.stack([(cols - cx) / fx, -(rows - cy) / fy, -torch.ones(H, W)], dim=-1)
    rot = torch.as_tensor(c2w, dtype=torch.float32)[:3, :3]
    d = torch.sum(d_cam[..., None, :] * rot, dim=-1).reshape(-1, 3)
    o = torch.as_tensor(c2w, dtype=torch.float32)[:3, 3].expand(d.shape)
    v = d / torch.norm(d, dim=-1, keepdim=True)
    nf = torch.tensor([near, far], dtype=torch.float32).expand(d.shape[0], 2)
    return torch.cat([o, d, nf, v], dim=-1)


# ---------------------------------------------------------------------------------------------
# R:357-379  coarse depths
# ---------------------------------------------------------------------------------------------
def coarse_depths(rays: Tensor, n_samples: int, lindisp: bool = False, t_rand: Optional[Tensor] = None) -> Tensor:
    near, far = rays[:, 6:7], rays[:, 7:8]
    t = torch.linspace(0.0, 1.0, steps=n_samples)
    if lindisp:
        z = 1.0 / (1.0 / near * (1.0 - t) + 1.0 / far * t)          # R:361
    else:
        z = near * (1.0 - t) + far * t                               # R:359
    z = z.expand(rays.shape[0], n_samples)
    if t_rand is not None:                                           # R:365-379
        mid = 0.5 * (z[:, 1:] + z[:, :-1])
        hi = torch.cat([mid, z[:, -1:]], dim=-1)
        lo = torch.cat([z[:, :1], mid], dim=-1)
        z = lo + (hi - lo) * t_rand
    return z


# ---------------------------------------------------------------------------------------------
# R:262-305  compositing
# ---------------------------------------------------------------------------------------------
def composite(raw: Tensor, z: Tensor, rays_d: Tensor, white_bkgd: bool = False, noise: Optional[Tensor] = None):
    """-> rgb_map, disp_map, acc_map, weights, depth_map.
    gaps = diff(z) with 1e10 appended, scaled by ||d|| (R:277-280); colour = sigmoid(raw[..., :3]) (R:282);
    alpha = 1 - exp(-relu(sigma + noise) * gap) (R:275,293); weights = alpha * exclusive cumprod(1 - alpha + 1e-10)
    (R:295); sums (R:296-300); disp = 1 / max(1e-10, depth / acc) (R:299); white background adds 1 - acc (R:303)."""
    gaps = torch.cat([z[:, 1:] - z[:, :-1], torch.full_like(z[:, :1], 1e10)], dim=-1)
    gaps = gaps * torch.norm(rays_d[:, None, :], dim=-1)
    colour = torch.sigmoid(raw[..., :3])
    sigma = raw[..., 3] if noise is None else raw[..., 3] + noise
    alpha = 1.0 - torch.exp(-F.relu(sigma) * gaps)
    trans = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1.0 - alpha + 1e-10], dim=-1), dim=-1)[:, :-1]
    w = alpha * trans
    rgb_map = torch.sum(w[..., None] * colour, dim=-2)
    depth_map = torch.sum(w * z, dim=-1)
    acc_map = torch.sum(w, dim=-1)
    disp_map = 1.0 / torch.max(1e-10 * torch.ones_like(depth_map), depth_map / acc_map)
    if white_bkgd:
        rgb_map = rgb_map + (1.0 - acc_map[..., None])
    return rgb_map, disp_map, acc_map, w, depth_map


# ---------------------------------------------------------------------------------------------
# H:200-243  inverse-CDF sampling
# ---------------------------------------------------------------------------------------------
def inverse_cdf_samples(bins: Tensor, weights: Tensor, n: int, u: Optional[Tensor] = None, return_inds: bool = False):
    """pdf from weights + 1e-5 (H:202-203); cdf = [0, cumsum(pdf)] (H:204-205); u = linspace(0,1,n) when
    deterministic (H:209-210); inds = searchsorted(cdf, u, right=True) (H:227); below = max(0, inds-1),
    above = min(len-1, inds) (H:228-229); denominators under 1e-5 become 1 (H:239); linear interpolation
    inside the bin (H:240-241)."""
    w = weights + 1e-5
    pdf = w / torch.sum(w, dim=-1, keepdim=True)
    cdf = torch.cat([torch.zeros_like(pdf[:, :1]), torch.cumsum(pdf, dim=-1)], dim=-1)
    if u is None:
        u = torch.linspace(0.0, 1.0, steps=n).expand(cdf.shape[0], n)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    lo = torch.clamp(inds - 1, min=0)
    hi = torch.clamp(inds, max=cdf.shape[-1] - 1)
    c_lo, c_hi = torch.gather(cdf, 1, lo), torch.gather(cdf, 1, hi)
    b_lo, b_hi = torch.gather(bins, 1, lo), torch.gather(bins, 1, hi)
    denom = c_hi - c_lo
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    out = b_lo + (u - c_lo) / denom * (b_hi - b_lo)
    return (out, inds) if return_inds else out


def hierarchical_depths(z_coarse: Tensor, weights: Tensor, n_importance: int, u: Optional[Tensor] = None):
    """R:392-396, R:412: bins are the coarse midpoints, the pdf uses weights[:, 1:-1], the new depths are merged
    with the coarse ones by a sort; z_std is the population std of the new depths."""
    mids = 0.5 * (z_coarse[:, 1:] + z_coarse[:, :-1])
    z_new = inverse_cdf_samples(mids, weights[:, 1:-1], n_importance, u).detach()
    z_all, _ = torch.sort(torch.cat([z_coarse, z_new], dim=-1), dim=-1)
    return z_all, z_new, torch.std(z_new, dim=-1, unbiased=False)


# ---------------------------------------------------------------------------------------------
# R:308-418 (+ C:418-421)  one ray batch, coarse + fine
# ---------------------------------------------------------------------------------------------
def render_ray_batch(rays: Tensor, sd_coarse, sd_fine, n_samples: int = 64, n_importance: int = 128,
                     white_bkgd: bool = True, lindisp: bool = False, t_rand: Optional[Tensor] = None,
                     u: Optional[Tensor] = None, retraw: bool = False):
    o, d, v = rays[:, 0:3], rays[:, 3:6], rays[:, 8:11]
    z = coarse_depths(rays, n_samples, lindisp, t_rand)
    pts = o[:, None, :] + d[:, None, :] * z[:, :, None]                              # R:381
    raw = query_network(sd_coarse, pts, v)
    rgb0, disp0, acc0, w0, _ = composite(raw, z, d, white_bkgd)
    out = {}
    if n_importance > 0:
        z, z_new, z_std = hierarchical_depths(z, w0, n_importance, u)
        pts = o[:, None, :] + d[:, None, :] * z[:, :, None]                          # R:397
        raw = query_network(sd_fine if sd_fine is not None else sd_coarse, pts, v)
        rgb, disp, acc, w, depth = composite(raw, z, d, white_bkgd)
        out.update(rgb0=rgb0, disp0=disp0, acc0=acc0, z_std=z_std)
    else:
        rgb, disp, acc, w, depth = rgb0, disp0, acc0, w0, None
    best = torch.argmax(w, dim=1)                                                     # C:418
    out.update(rgb_map=rgb, disp_map=disp, acc_map=acc, pts_max=pts[torch.arange(pts.shape[0]), best],  # C:421
               weights=w, z_vals=z)
    if retraw:
        out["raw"] = raw
    return out


def render_image(H, W, K, c2w, sd_coarse, sd_fine, near=2.0, far=6.0, chunk=1024, **kw):
    """render (R:69-134) for a full image: rays from the camera, chunked (R:54-66), outputs reshaped to [H,W,...]."""
    rays = camera_rays(H, W, K, c2w, near, far)
    parts = [render_ray_batch(rays[i:i + chunk], sd_coarse, sd_fine, **kw) for i in range(0, rays.shape[0], chunk)]
    return {k: torch.cat([p[k] for p in parts], dim=0).reshape(H, W, *parts[0][k].shape[1:]) for k in parts[0]}


# ---------------------------------------------------------------------------------------------
# bf16 emulation of the fused kernel's numerics (test infrastructure for nerfail_b200/csrc/mlp_fused.cu)
# ---------------------------------------------------------------------------------------------
def _bf16(t: Tensor) -> Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def nerf_mlp_bf16_emulated(sd: Dict[str, Tensor], pts: Tensor, dirs: Tensor, return_steps: bool = False):
    """Same network as nerf_mlp (H:100-121) with the rounding points of the tcgen05 kernel: encodings, weights and
    the activations handed from layer to layer are bf16; products are exact and sums fp32; biases, the sigma head
    (on the un-rounded layer-7 activations) and the rgb head stay fp32.  pts [M,3], dirs [M,3] -> [M,4]."""
    mm = lambda a, w: a.double().matmul(_bf16(w).double().t()).float()
    e_pts = _bf16(positional_encoding(pts, 10))
    e_dir = _bf16(positional_encoding(dirs, 4))
    steps = []
    h = e_pts
    h_f32 = None
    for i in range(8):
        h_f32 = F.relu(mm(h, sd[f"pts_linears.{i}.weight"]) + sd[f"pts_linears.{i}.bias"])
        steps.append(h_f32)
        h = _bf16(h_f32)
        if i == 4:
            h = torch.cat([e_pts, h], dim=-1)
    sigma = F.linear(h_f32, sd["alpha_linear.weight"], sd["alpha_linear.bias"])
    feat_f32 = mm(h, sd["feature_linear.weight"]) + sd["feature_linear.bias"]
    steps.append(feat_f32)
    hv = F.relu(mm(torch.cat([_bf16(feat_f32), e_dir], dim=-1), sd["views_linears.0.weight"]) + sd["views_linears.0.bias"])
    steps.append(hv)
    rgb = F.linear(hv, sd["rgb_linear.weight"], sd["rgb_linear.bias"])
    out = torch.cat([rgb, sigma], dim=-1)
    return (out, steps) if return_steps else out


# ---------------------------------------------------------------------------------------------
# R:744-773  ray-batch sampling of the no_batching path (the only one NeRFail's configs use)
# ---------------------------------------------------------------------------------------------
def sample_ray_batch(images, poses, i_train, H: int, W: int, K, N_rand: int, step: int, precrop_iters: int,
                     precrop_frac: float, rng=np.random):
    """A random training image (R:746), all of its rays (H:157-166 via camera_rays), the pixel grid or its centre crop
    for the first precrop_iters iterations (R:754-766), N_rand distinct pixels (R:768) -> batch_rays [2,N,3],
    target_s [N,3], img_i, select_coords [N,2]."""
    img_i = int(rng.choice(np.asarray(i_train)))
    target = torch.as_tensor(np.asarray(images[img_i]), dtype=torch.float32)
    pose = torch.as_tensor(np.asarray(poses[img_i]), dtype=torch.float32)[:3, :4]
    rays = camera_rays(H, W, K, pose, 0.0, 1.0)
    rays_o, rays_d = rays[:, 0:3].reshape(H, W, 3), rays[:, 3:6].reshape(H, W, 3)
    if step < precrop_iters:
        dH = int(H // 2 * precrop_frac)
        dW = int(W // 2 * precrop_frac)
        coords = torch.stack(torch.meshgrid(torch.linspace(H // 2 - dH, H // 2 + dH - 1, 2 * dH),
                                            torch.linspace(W // 2 - dW, W // 2 + dW - 1, 2 * dW), indexing="ij"), -1)
    else:
        coords = torch.stack(torch.meshgrid(torch.linspace(0, H - 1, H), torch.linspace(0, W - 1, W), indexing="ij"), -1)
    coords = torch.reshape(coords, [-1, 2])
    select_inds = rng.choice(coords.shape[0], size=[N_rand], replace=False)
    select_coords = coords[select_inds].long()
    rays_o = rays_o[select_coords[:, 0], select_coords[:, 1]]
    rays_d = rays_d[select_coords[:, 0], select_coords[:, 1]]
    target_s = target[select_coords[:, 0], select_coords[:, 1]][..., :3]
    return torch.stack([rays_o, rays_d], 0), target_s, img_i, select_coords


# ---------------------------------------------------------------------------------------------
# R:776-800  the core optimisation loop (loss, backward, Adam, exponential learning-rate decay)
# ---------------------------------------------------------------------------------------------
class Trainer:
    """The reference's optimisation step on the CPU, functional weights: render the batch (R:776-778), loss =
    img2mse(fine) + img2mse(coarse) (R:781-789, H:9), backward, torch.optim.Adam(lr, betas=(0.9, 0.999)) exactly as
    create_nerf builds it (R:213), then lr = lrate * 0.1 ** (global_step / (lrate_decay * 1000)) (R:796-800).
    Pinned by tests/golden/train_traj.npz (three steps of the unmodified reference, make_golden_train.py)."""

    def __init__(self, sd_coarse, sd_fine, lrate: float = 5e-4, lrate_decay: int = 250):
        self.sd_c = {k: v.clone().float().requires_grad_(True) for k, v in sd_coarse.items()}
        self.sd_f = {k: v.clone().float().requires_grad_(True) for k, v in sd_fine.items()}
        self.lrate, self.lrate_decay = lrate, lrate_decay
        self.opt = torch.optim.Adam(list(self.sd_c.values()) + list(self.sd_f.values()), lr=lrate, betas=(0.9, 0.999))

    def loss_and_outputs(self, rays: Tensor, target: Tensor, t_rand: Optional[Tensor] = None, u: Optional[Tensor] = None,
                         n_samples: int = 64, n_importance: int = 128, white_bkgd: bool = True):
        out = render_ray_batch(rays, self.sd_c, self.sd_f, n_samples, n_importance, white_bkgd, False, t_rand, u)
        loss = torch.mean((out["rgb_map"] - target) ** 2) + torch.mean((out["rgb0"] - target) ** 2)
        return loss, out

    def step(self, rays: Tensor, target: Tensor, global_step: int, t_rand: Optional[Tensor] = None, u: Optional[Tensor] = None,
             **kw) -> float:
        self.opt.zero_grad()
        loss, _ = self.loss_and_outputs(rays, target, t_rand, u, **kw)
        loss.backward()
        self.opt.step()
        new_lrate = self.lrate * (0.1 ** (global_step / (self.lrate_decay * 1000)))
        for gp in self.opt.param_groups:
            gp["lr"] = new_lrate
        return float(loss)

    def state_dicts(self):
        return ({k: v.detach().clone() for k, v in self.sd_c.items()}, {k: v.detach().clone() for k, v in self.sd_f.items()})


def rays_from_batch(batch_rays: Tensor, near: float, far: float) -> Tensor:
    """render(rays=batch_rays) of R:95-123 for use_viewdirs=True, ndc=False: [2,N,3] -> the [N,11] ray batch."""
    o, d = batch_rays[0].float(), batch_rays[1].float()
    v = d / torch.norm(d, dim=-1, keepdim=True)
    n = torch.ones_like(d[:, :1])
    return torch.cat([o, d, near * n, far * n, v], -1)
