"""The NeRF (re)training step around the render kernels: batch sampler, optimisation step, learning-rate decay and
checkpoints, with the reference's semantics and file format.

Reference: Create_spatial_point_set/nerf_pytorch/run_nerf.py — ray-batch sampling of the `no_batching` path (:744-773,
what every NeRFail config uses, configs/lego.txt), the core optimisation loop (:776-801) and the checkpoint written
every i_weights iterations (:808-816, re-read by create_nerf :216-233).  The attack retrains NeRF on perturbed images
with exactly this loop (`run_nerf.py --train_dir`, README.md:187-205).

Data-parallel form: every rank draws ITS rays of the step's batch (sample_ray_batch with rank /
world_size takes the rank's contiguous share of the same N_rand pixels), the mean-squared losses are means over the
rank's rays, so gradients are all-reduced with scale 1/G (one flat bucket per network) and every rank applies the
identical fused Adam step.
"""
from __future__ import annotations

import os
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import dist as nd
from . import nerf
from .optim import decayed_lrate, set_lrate
from .rendering import render


def precrop_window(H: int, W: int, precrop_frac: float) -> Tuple[int, int, int, int]:
    """(row0, col0, rows, cols) of the centre crop used for the first precrop_iters iterations (run_nerf.py:754-763)."""
    dH = int(H // 2 * precrop_frac)
    dW = int(W // 2 * precrop_frac)
    return H // 2 - dH, W // 2 - dW, 2 * dH, 2 * dW


def sample_ray_batch(images, poses, i_train: Sequence[int], H: int, W: int, K, N_rand: int, step: int,
                     precrop_iters: int = 0, precrop_frac: float = 0.5, rng=np.random, device=None,
                     rank: int = 0, world_size: int = 1):
    """One `no_batching` batch (run_nerf.py:744-773): a random training image, N_rand distinct pixels of it (of its centre
    crop while step < precrop_iters), their rays and target colours.  Consumes the host RNG exactly like the reference
    (`choice(i_train)` then `choice(n_pixels, size=[N_rand], replace=False)`), so a seeded run picks the same pixels.
    Only the selected pixels' rays are formed (the reference builds all H*W rays and indexes them), on the device.
    images: [N,H,W,3] array or tensor (host or device), poses: [N,>=3,4].  Returns (batch_rays [2,n,3], target_s [n,3],
    img_i, select_coords [n,2] (row, col) int64) where n = this rank's share of N_rand."""
    device = torch.device(device if device is not None else "cuda")
    img_i = int(rng.choice(np.asarray(i_train)))
    if step < precrop_iters:
        r0, c0, nr, nc = precrop_window(H, W, precrop_frac)
    else:
        r0, c0, nr, nc = 0, 0, H, W
    select_inds = rng.choice(nr * nc, size=[N_rand], replace=False)
    b, e = nd.shard_range(N_rand, rank, world_size)
    sel = torch.from_numpy(np.asarray(select_inds[b:e], dtype=np.int64)).to(device)
    rows = r0 + sel // nc                      # coords = meshgrid(linspace(rows), linspace(cols)) flattened row-major (:756-766)
    cols = c0 + sel % nc
    pose = torch.as_tensor(np.asarray(poses[img_i])[:3, :4] if not isinstance(poses, torch.Tensor) else poses[img_i, :3, :4],
                           dtype=torch.float32).to(device)
    # get_rays (run_nerf_helpers.py:157-166) restricted to the selected pixels: i = column, j = row
    fi, fj = cols.to(torch.float32), rows.to(torch.float32)
    dirs = torch.stack([(fi - K[0][2]) / K[0][0], -(fj - K[1][2]) / K[1][1], -torch.ones_like(fi)], -1)
    rays_d = torch.sum(dirs[..., None, :] * pose[:3, :3], -1)
    rays_o = pose[:3, -1].expand(rays_d.shape)
    img = images[img_i]
    if isinstance(img, torch.Tensor):
        target_s = img.to(device)[rows, cols]
    else:                                      # host image: move only the N_rand selected pixels
        target_s = torch.from_numpy(np.asarray(img)[rows.cpu().numpy(), cols.cpu().numpy()]).to(device)
    target_s = target_s[..., :3].to(torch.float32)
    return torch.stack([rays_o, rays_d], 0), target_s, img_i, torch.stack([rows, cols], -1)


def train_step(batch_rays, target_s, H: int, W: int, K, chunk: int, render_kwargs_train: dict, optimizer,
               lrate: float, lrate_decay: int, global_step: int, near: Optional[float] = None,
               far: Optional[float] = None) -> dict:
    """One optimisation step (run_nerf.py:776-800): render the batch with retraw, loss = mse(fine) + mse(coarse),
    backward, (data-parallel: all-reduce of the gradients, scale 1/G), optimizer.step(), exponential lr decay evaluated
    at global_step like the reference.  Returns {'loss', 'psnr', 'psnr0'} as device scalars (no host sync)."""
    kw = dict(render_kwargs_train)
    if near is not None:
        kw.update(near=near, far=far)
    with torch.enable_grad():
        rgb, disp, acc, extras = render(H, W, K, chunk=chunk, rays=batch_rays, retraw=True, **kw)
        optimizer.zero_grad()
        img_loss = nerf.img2mse(rgb, target_s)
        loss = img_loss
        out = {"psnr": nerf.mse2psnr(img_loss.detach())}
        if "rgb0" in extras:
            img_loss0 = nerf.img2mse(extras["rgb0"], target_s)
            loss = loss + img_loss0
            out["psnr0"] = nerf.mse2psnr(img_loss0.detach())
        loss.backward()
    _, world_size = nd.world()
    if world_size > 1:
        nets = [kw.get("network_fn"), kw.get("network_fine")]
        pending = [nd.allreduce_grads_(n.parameters(), scale=1.0 / world_size, async_op=True) for n in nets if n is not None]
        for fin in pending:
            if fin is not None:
                fin()
    optimizer.step()
    set_lrate(optimizer, decayed_lrate(lrate, lrate_decay, global_step))
    out["loss"] = loss.detach()
    return out


def save_checkpoint(path: str, global_step: int, render_kwargs_train: dict, optimizer) -> str:
    """The .tar of run_nerf.py:808-816 (same four keys), readable by the reference's and this package's create_nerf."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    ckpt = {
        "global_step": global_step,
        "network_fn_state_dict": render_kwargs_train["network_fn"].state_dict(),
        "optimizer_state_dict": optimizer.state_dict(),
    }
    if render_kwargs_train.get("network_fine") is not None:
        ckpt["network_fine_state_dict"] = render_kwargs_train["network_fine"].state_dict()
    torch.save(ckpt, path)
    return path


def checkpoint_path(basedir: str, expname: str, i: int) -> str:
    return os.path.join(basedir, expname, "{:06d}.tar".format(i))
