"""Oracle restatement of the GaussNet path and the 8-NN precompute (TEST INFRASTRUCTURE — see oracle/__init__.py).

Citations relative to the reference tree:  G = model/GaussNet.py,  I = Create_spatial_point_set/create_index_and_dist.py
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

Tensor = torch.Tensor


def gaussian_weights(dist_idx: Tensor, c: float) -> Tensor:
    """create_gauss_w.forward (G:169-186): d = exp(-(dist/c)^2 / 2); w = d / (sum_8 d + 0.001) where the sum is
    positive, else 0; returns cat([w, idx]) on axis 1 -> [B,2,H,W,8]."""
    dist, idx = dist_idx[:, 0:1], dist_idx[:, 1:2]
    d = torch.exp(-(torch.square(dist / c) / 2))
    s = torch.sum(d, dim=-1, keepdim=True).expand_as(d)
    w = torch.where(s > 0, d / (s + 0.001), torch.zeros_like(d))
    return torch.cat([w, idx], dim=1)


def gauss_forward(spatial_rgb: Tensor, w_idx: Tensor, ori_img_u8: Tensor, epsilon: Optional[float] = None):
    """gauss_net.forward up to x_rgba (G:53-119).  Returns (x [B,H,W,4], x_rgba [B,H,W,4], (eps_min, eps_max)).
    x = sum_k w_k * table[idx_k] (G:53-83); alpha = x_a / 255 (G:85); the tracked extrema are those of
    alpha * where(alpha > 0, x_rgb, 0) (G:91-97); rgb = ori_rgb + clip(x_rgb * alpha, +-eps) (G:106-110), zeroed
    where the original alpha is 0 (G:112-113); x_rgba = clip(cat(rgb, ori_alpha), 0, 255) (G:116-119)."""
    table = spatial_rgb.reshape(-1, 4)
    ori = ori_img_u8.to(torch.float32)
    w, idx = w_idx[:, 0], w_idx[:, 1].to(torch.long)
    B, H, W_, Kn = idx.shape
    rows = table[idx.reshape(-1)].reshape(B, H, W_, Kn, 4)
    x = torch.sum(rows * w[..., None], dim=-2)
    alpha = x[..., 3:4] / 255
    masked = torch.where(alpha.expand_as(x[..., :3]) > 0, x[..., :3], torch.zeros_like(x[..., :3])) * alpha
    extrema = (float(masked.detach().min()), float(masked.detach().max()))
    delta = x[..., :3] * alpha
    if epsilon is not None:
        delta = torch.clip(delta, -epsilon, epsilon)
    rgb = torch.where(ori[..., 3:4] > 0, ori[..., :3] + delta, torch.zeros_like(delta))
    x_rgba = torch.clip(torch.cat([rgb, ori[..., 3:4]], dim=-1), min=0, max=255)
    return x, x_rgba, extrema


def knn8_exact(query: np.ndarray, cand: np.ndarray, block: int = 2048):
    """8 nearest candidates per query with the semantics of I:126-145 (Euclidean distance, ascending, global
    candidate index) but in the exact direct-difference form SURVEY.md §0.4/§8c prescribes as the parity anchor:
    d2 = ((dx*dx + dy*dy) + dz*dz) in fp32 with no fused multiply-add, ordered by (d2, index); dist = sqrt(d2).
    Returns (dist [Q,8] float32, idx [Q,8] int32)."""
    q = np.ascontiguousarray(query, dtype=np.float32).reshape(-1, 3)
    c = np.ascontiguousarray(cand, dtype=np.float32).reshape(-1, 3)
    out_d = np.empty((q.shape[0], 8), np.float32)
    out_i = np.empty((q.shape[0], 8), np.int32)
    for s in range(0, q.shape[0], block):
        qq = q[s:s + block]
        dx = qq[:, None, 0] - c[None, :, 0]
        dy = qq[:, None, 1] - c[None, :, 1]
        dz = qq[:, None, 2] - c[None, :, 2]
        d2 = (dx * dx + dy * dy) + dz * dz                     # numpy evaluates each product/sum separately in fp32
        order = np.argsort(d2, axis=1, kind="stable")[:, :8]   # stable: equal distances keep the lower index first
        out_i[s:s + block] = order.astype(np.int32)
        out_d[s:s + block] = np.sqrt(np.take_along_axis(d2, order, axis=1))
    return out_d, out_i


def knn8_reference_style(query: Tensor, cand: Tensor, chunk: int = 1200, compute_mode: str = "use_mm_for_euclid_dist_if_necessary"):
    """The reference's own procedure (I:126-145): cdist against candidate chunks, sort, keep 8, merge with the
    running 8 by cat + sort + gather.  Used only to REPORT agreement statistics (its matmul-mode cdist is
    numerically noisy)."""
    best_d = best_i = None
    offset = 0
    for part in torch.split(cand, chunk, dim=0):
        d = torch.cdist(query, part, compute_mode=compute_mode)
        vals, ids = torch.sort(d, dim=-1)
        vals, ids = vals[..., :8], ids[..., :8] + offset
        offset += part.shape[0]
        if best_d is not None:
            vals, ids = torch.cat([best_d, vals], dim=-1), torch.cat([best_i, ids], dim=-1)
            vals, order = torch.sort(vals, dim=-1)
            ids = ids.gather(dim=-1, index=order)
        best_d, best_i = vals[..., :8], ids[..., :8]
    return best_d, best_i
