"""Seeded synthetic inputs shared by the oracle, the tests and bench.py.

Generators only — no reference arithmetic here except pose_spherical, which restates
Create_spatial_point_set/nerf_pytorch/load_blender.py:11-34 (Blender-style cameras on a sphere).
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch

CAMERA_ANGLE_X = 0.6911112070083618      # nerf_synthetic lego field of view (any fixed value works)


def pose_spherical(theta_deg: float, phi_deg: float, radius: float) -> np.ndarray:
    """load_blender.py:29-34: translate along z, rotate by phi about x, by theta about y, then the fixed
    axis flip [[-1,0,0,0],[0,0,1,0],[0,1,0,0],[0,0,0,1]].  Returns a 4x4 float32 camera-to-world matrix."""
    t = np.eye(4, dtype=np.float32); t[2, 3] = radius
    p, th = np.deg2rad(phi_deg), np.deg2rad(theta_deg)
    rx = np.array([[1, 0, 0, 0], [0, np.cos(p), -np.sin(p), 0], [0, np.sin(p), np.cos(p), 0], [0, 0, 0, 1]], np.float32)
    ry = np.array([[np.cos(th), 0, -np.sin(th), 0], [0, 1, 0, 0], [np.sin(th), 0, np.cos(th), 0], [0, 0, 0, 1]], np.float32)
    flip = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], np.float32)
    return (flip @ ry @ rx @ t).astype(np.float32)


def intrinsics(H: int, W: int):
    """load_blender.py:84-85 focal; run_nerf.py:631-636 K."""
    focal = 0.5 * W / np.tan(0.5 * CAMERA_ANGLE_X)
    return np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]]), focal


def camera_ring(n_views: int):
    return [pose_spherical(a, -30.0, 4.0) for a in np.linspace(-180, 180, n_views + 1)[:-1]]


# --- network weights ------------------------------------------------------------------------------
_SHAPES = OrderedDict(
    [(f"pts_linears.{i}", (256, 63 if i == 0 else (319 if i == 5 else 256))) for i in range(8)]
    + [("views_linears.0", (128, 283)), ("feature_linear", (256, 256)), ("alpha_linear", (1, 256)),
       ("rgb_linear", (3, 128))])


def random_state_dict(seed: int) -> "OrderedDict[str, torch.Tensor]":
    """W-A: nn.Linear default init (uniform +-1/sqrt(fan_in) for weight and bias) of the lego architecture,
    in nn.Module registration order."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for name, (n_out, n_in) in _SHAPES.items():
        bound = 1.0 / np.sqrt(n_in)
        sd[f"{name}.weight"] = (torch.rand(n_out, n_in, generator=g) * 2 - 1) * bound
        sd[f"{name}.bias"] = (torch.rand(n_out, generator=g) * 2 - 1) * bound
    return sd


def make_non_degenerate(sd, seed: int, target_std: float = 2.0, rgb_gain: float = 20.0):
    """W-B: affine-normalise the sigma head on 8192 probe points so that density varies in space
    (random init gives sigma of one sign everywhere), and widen the rgb logits."""
    from . import nerf_oracle as no
    g = torch.Generator().manual_seed(1000 + seed)
    p = (torch.rand(8192, 3, generator=g) * 3 - 1.5)
    v = torch.randn(8192, 3, generator=g)
    v = v / v.norm(dim=-1, keepdim=True)
    with torch.no_grad():
        sig = no.query_network(sd, p[:, None, :], v)[:, 0, 3]
        k = target_std / sig.std()
        out = OrderedDict((n, t.clone()) for n, t in sd.items())
        out["alpha_linear.weight"] = sd["alpha_linear.weight"] * k
        out["alpha_linear.bias"] = k * sd["alpha_linear.bias"] - k * sig.median()
        out["rgb_linear.weight"] = sd["rgb_linear.weight"] * rgb_gain
    return out


def flat_params(sd) -> torch.Tensor:
    """state_dict order expected by nfb_mlp_update."""
    order = [f"pts_linears.{i}" for i in range(8)] + ["views_linears.0", "feature_linear", "alpha_linear", "rgb_linear"]
    return torch.cat([torch.cat([sd[f"{n}.weight"].reshape(-1), sd[f"{n}.bias"].reshape(-1)]) for n in order])


# --- GaussNet inputs ------------------------------------------------------------------------------
def gauss_inputs(seed: int, P: int = 3, H: int = 64, W: int = 64, B: int = 2, locality: bool = True):
    """spatial_rgb [P,H,W,4] (RGB ~ N(0,5^2), A = 255 on a centred disc else 0), dist/idx [B,2,H,W,8]
    (ascending |N(0,0.01^2)| distances; indices either near a per-pixel anchor or uniform), ori uint8 [B,H,W,4]."""
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    disc = (((yy - H / 2) ** 2 + (xx - W / 2) ** 2) < (0.36 * min(H, W)) ** 2).float()
    s = torch.randn(P, H, W, 4, generator=g) * 5.0
    s[..., 3] = disc * 255.0
    T = P * H * W
    dist = torch.sort(torch.randn(B, H, W, 8, generator=g).abs() * 0.01, dim=-1).values
    if locality:
        anchor = (torch.arange(H * W).reshape(1, H, W, 1) + torch.randint(0, P, (B, 1, 1, 1), generator=g) * H * W)
        idx = (anchor + torch.randint(-3, 4, (B, H, W, 8), generator=g)).clamp(0, T - 1)
    else:
        idx = torch.randint(0, T, (B, H, W, 8), generator=g)
    dist_idx = torch.stack([dist, idx.float()], dim=1)
    ori = torch.randint(0, 256, (B, H, W, 4), generator=g).to(torch.uint8)
    ori[..., 3] = (disc * 255).to(torch.uint8)
    return s, dist_idx, ori


def write_blender_scene(root: str, H: int = 8, W: int = 8, counts=(3, 2, 4), seed: int = 0, train_dir: str = None):
    """A tiny nerf_synthetic-style scene on disk for the loader tests: transforms_{train,val,test}.json with
    `camera_angle_x` and per-frame `file_path` / `transform_matrix` (the fields load_blender.py:37-85 reads), RGBA PNGs
    with a transparent border.  train_dir (optional) receives a second, different set of train images under the same
    base names — what `run_nerf.py --train_dir` points at after an attack (load_blender.py:62-63)."""
    import json
    import os

    import cv2
    rng = np.random.default_rng(seed)
    th = 0
    for split, n in zip(("train", "val", "test"), counts):
        os.makedirs(os.path.join(root, split), exist_ok=True)
        frames = []
        for i in range(n):
            img = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
            img[0, :, 3] = 0
            img[:, -1, 3] = 0
            cv2.imwrite(os.path.join(root, split, f"r_{i}.png"), img[..., [2, 1, 0, 3]])       # cv2 writes BGRA
            if split == "train" and train_dir is not None:
                os.makedirs(train_dir, exist_ok=True)
                att = np.clip(img.astype(np.int32) + rng.integers(-8, 9, img.shape), 0, 255).astype(np.uint8)
                att[..., 3] = img[..., 3]
                cv2.imwrite(os.path.join(train_dir, f"r_{i}.png"), att[..., [2, 1, 0, 3]])
            frames.append({"file_path": f"./{split}/r_{i}", "rotation": 0.1,
                           "transform_matrix": pose_spherical(-180.0 + 37.0 * th, -30.0, 4.0).tolist()})
            th += 1
        with open(os.path.join(root, f"transforms_{split}.json"), "w") as fp:
            json.dump({"camera_angle_x": 0.6911112070083618, "frames": frames}, fp)
    return root
