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
from . import nerf, ops
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


def _poll_networks(kw: dict) -> None:
    """Raises if a fused kernel of either network reported a pipeline-barrier time-out so far (no synchronisation: one
    read of pinned host memory per network).  A step whose kernels are still in flight is covered by the next call."""
    for key in ("network_fn", "network_fine"):
        n = kw.get(key)
        if isinstance(n, nerf.NeRF) and n._fused is not None and not torch.cuda.is_current_stream_capturing():
            n._fused.poll()


def train_step(batch_rays, target_s, H: int, W: int, K, chunk: int, render_kwargs_train: dict, optimizer,
               lrate: float, lrate_decay: int, global_step: int, near: Optional[float] = None,
               far: Optional[float] = None, exchange=None, _begun: bool = False) -> dict:
    """One optimisation step (run_nerf.py:776-800): render the batch with retraw, loss = mse(fine) + mse(coarse),
    backward, (data-parallel: all-reduce of the gradients, scale 1/G), optimizer.step(), exponential lr decay evaluated
    at global_step like the reference.  Returns {'loss', 'psnr', 'psnr0'} as device scalars (no host sync).
    exchange: a dist.PeerAdam over the same networks / optimizer — the gradient average, Adam and the parameter broadcast
    then are one kernel per GPU over NVLink peer memory instead of two NCCL all-reduces and a replicated Adam."""
    kw = dict(render_kwargs_train)
    if near is not None:
        kw.update(near=near, far=far)
    with torch.enable_grad():
        rgb, disp, acc, extras = render(H, W, K, chunk=chunk, rays=batch_rays, retraw=True, **kw)
        optimizer.zero_grad()
        # img2mse(fine) + img2mse(coarse) and their gradients in one kernel (run_nerf.py:781-789)
        loss, _mse, psnr = ops.MseLoss2Fn.apply(rgb, extras.get("rgb0"), target_s[..., :3])
        out = {"psnr": psnr[0]}
        if "rgb0" in extras:
            out["psnr0"] = psnr[1]
        loss.backward()
    _, world_size = nd.world()
    if exchange is not None:
        exchange.step(begin=not _begun)
    else:
        if world_size > 1:
            nets = [kw.get("network_fn"), kw.get("network_fine")]
            pending = [nd.allreduce_grads_(n.parameters(), scale=1.0 / world_size, async_op=True) for n in nets if n is not None]
            for fin in pending:
                if fin is not None:
                    fin()
        optimizer.step()
    set_lrate(optimizer, decayed_lrate(lrate, lrate_decay, global_step))
    _poll_networks(kw)
    out["loss"] = loss.detach()
    return out


class GraphedTrainStep:
    """train_step captured ONCE in a CUDA graph and replayed: the ~60 launches of a step (ray setup, two fused forward
    passes, compositing, hierarchical sampling, loss, compositing / data-gradient / weight-gradient kernels, gradient
    all-reduce, fused Adam, weight re-pack) cost one graph launch on the host instead of ~2.6 ms of Python, which is what
    bounds a data-parallel step once a rank's share of the batch is below ~2000 rays.  Everything that changes from
    step to step lives in device memory the graph reads: the ray batch and targets (static buffers filled before the
    replay), the sampler's random numbers (torch's graph-safe Philox state), Adam's step size and bias correction
    (optim.Adam.begin_step).  Same arithmetic as train_step; checked by test_graphed_train_step_equals_eager."""

    def __init__(self, n_rays: int, H: int, W: int, K, chunk: int, render_kwargs_train: dict, optimizer, lrate: float,
                 lrate_decay: int, near: Optional[float] = None, far: Optional[float] = None, device=None, warmup: int = 3,
                 exchange=None):
        self.device = torch.device(device if device is not None else "cuda")
        self.args = (H, W, K, chunk)
        self.kw, self.optimizer = render_kwargs_train, optimizer
        self.lrate, self.lrate_decay, self.near, self.far = lrate, lrate_decay, near, far
        self.batch_rays = torch.zeros(2, n_rays, 3, device=self.device)
        self.target_s = torch.zeros(n_rays, 3, device=self.device)
        self.warmup = warmup
        self.exchange = exchange           # dist.PeerAdam: fused average + Adam + broadcast over NVLink peer memory
        self.graph = None
        self.out = None
        optimizer.enable_graph_mode(self.device)

    def _nets(self):
        return [n for n in (self.kw.get("network_fn"), self.kw.get("network_fine")) if n is not None]

    def _eager(self, global_step):
        H, W, K, chunk = self.args
        return train_step(self.batch_rays, self.target_s, H, W, K, chunk, self.kw, self.optimizer, self.lrate,
                          self.lrate_decay, global_step, self.near, self.far, exchange=self.exchange, _begun=True)

    def __call__(self, batch_rays, target_s, global_step: int) -> dict:
        """One step on (batch_rays [2,n,3], target_s [n,3]); returns {'loss', 'psnr', 'psnr0'} (device scalars that the
        next call overwrites).  The first `warmup` calls run eagerly on a side stream (they are real steps), the next one
        is captured, every later one is a replay."""
        self.batch_rays.copy_(batch_rays, non_blocking=True)
        self.target_s.copy_(target_s, non_blocking=True)
        self.optimizer.begin_step()
        if self.graph is not None:
            self.graph.replay()
        elif self.warmup > 0:
            self.warmup -= 1
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                self.out = self._eager(global_step)
            torch.cuda.current_stream(self.device).wait_stream(side)
        else:
            self.optimizer.zero_grad(set_to_none=True)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.out = self._eager(global_step)
            self.graph = g
            g.replay()                      # capture does not execute: this replay IS the step
        # the graph's Adam kernel moved the parameters behind torch's back: force the next non-graph user to re-pack
        for n in self._nets():
            n.invalidate_fused()
        _poll_networks(self.kw)
        set_lrate(self.optimizer, decayed_lrate(self.lrate, self.lrate_decay, global_step))
        return self.out


def make_train_stepper(n_rays: int, H: int, W: int, K, chunk: int, render_kwargs_train: dict, optimizer, lrate: float,
                       lrate_decay: int, near: Optional[float] = None, far: Optional[float] = None, device=None,
                       graph: bool = True, peer: Optional[bool] = None):
    """The recommended way to run the optimisation loop of run_nerf.py:776-800: returns `step(batch_rays, target_s, global_step)
    -> {'loss', 'psnr', 'psnr0'}` and the PeerAdam object (or None).
    Under torch.distributed (world size > 1, one process per GPU of a node) the gradient average, Adam and the parameter
    broadcast are the fused peer-memory kernel (dist.PeerAdam) unless peer=False keeps the NCCL all-reduces; graph=True
    captures the whole step in one CUDA graph (GraphedTrainStep), which is what makes per-rank batches of a few hundred
    rays worth sharding at all.  n_rays = this rank's share of the batch (sample_ray_batch(..., rank, world_size))."""
    _, world_size = nd.world()
    use_peer = (world_size > 1) if peer is None else bool(peer)
    exchange = None
    if use_peer:
        nets = [render_kwargs_train.get("network_fn"), render_kwargs_train.get("network_fine")]
        exchange = nd.PeerAdam(nets, optimizer, device)
    if graph:
        stepper = GraphedTrainStep(n_rays, H, W, K, chunk, render_kwargs_train, optimizer, lrate, lrate_decay, near, far,
                                   device=device, exchange=exchange)
        return stepper, exchange

    def step(batch_rays, target_s, global_step):
        return train_step(batch_rays, target_s, H, W, K, chunk, render_kwargs_train, optimizer, lrate, lrate_decay, global_step,
                          near, far, exchange=exchange)
    return step, exchange


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
