"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the GPU box, gloo in CPU tests).

The reference is single-process; what shards and what must be exchanged follows SURVEY.md §8e:
  * render sweeps (run_nerf.py:151-154): views are independent -> shard by view, no collective;
  * one view: rays are independent -> contiguous ray ranges, no collective;
  * NeRFail-S attack iteration (attack_NeRFail_S.py:304-392): perturbation replicated, views sharded, ONE all-reduce
    (sum) of grad_spatial_rgb [P,H,W,4] before the sign step, so every rank applies the identical update;
  * NeRF retraining (run_nerf.py:776-792): rays sharded, ONE all-reduce of the parameter gradients per network.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_views(n_views: int, rank: int, world_size: int) -> List[int]:
    """View i -> rank i mod G (BASELINE config 4)."""
    return list(range(rank, n_views, world_size))


def shard_range(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) of n items for this rank (ray ranges of one view / of one training batch)."""
    base, rem = divmod(n, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def allreduce_sum_(t: torch.Tensor, async_op: bool = False):
    """In-place sum over ranks; no-op handle when not distributed."""
    if world()[1] == 1:
        return None
    return dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=async_op)


def _flat_view(grads):
    """The gradients as ONE tensor without a copy, if they are back-to-back views of one buffer (what the fused training
    backward hands to autograd: nfb_mlp_bwd_weights accumulates into a flat [n_params] gradient in state_dict order)."""
    g0 = grads[0]
    base, off = g0.untyped_storage().data_ptr(), g0.storage_offset()
    for g in grads:
        if g.untyped_storage().data_ptr() != base or g.dtype != g0.dtype or not g.is_contiguous() or g.storage_offset() != off:
            return None
        off += g.numel()
    return torch.as_strided(g0, (off - g0.storage_offset(),), (1,), g0.storage_offset())


def allreduce_grads_(params: Iterable[torch.nn.Parameter], scale: float = 1.0, async_op: bool = False):
    """One flat bucket per call (4.77 MB for both NeRF networks): all-reduce, scale.  Gradients that already live in
    one flat buffer are reduced in place (no flatten / scatter kernels); otherwise flatten, all-reduce, scatter back.
    Call once per network so the coarse bucket is in flight while the fine network's wgrad still runs."""
    ps = [p for p in params if p.grad is not None]
    if not ps:
        return None
    flat = _flat_view([p.grad for p in ps])
    in_place = flat is not None
    if not in_place:
        flat = torch.cat([p.grad.reshape(-1) for p in ps])
    work = allreduce_sum_(flat, async_op=async_op)

    def finish():
        if work is not None and async_op:
            work.wait()
        if scale != 1.0:
            flat.mul_(scale)
        if in_place:
            return
        off = 0
        for p in ps:
            n = p.grad.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n
    if async_op:
        return finish
    finish()
    return None


def attack_sign_step_(spatial_rgb: torch.Tensor, grad: torch.Tensor, init: torch.Tensor, step: float, eps: float,
                      minimise: bool = True) -> torch.Tensor:
    """The I-FGSM update of attack_NeRFail_S.py:357-392 applied AFTER the gradient all-reduce: sign step on the RGB
    channels where the point is active (A > 0), then clamp to init +- eps.  Identical on every rank by construction."""
    allreduce_sum_(grad)
    active = (spatial_rgb[..., 3:4] > 0).to(spatial_rgb.dtype)
    delta = step * torch.sign(grad[..., :3]) * active
    rgb = spatial_rgb[..., :3] - delta if minimise else spatial_rgb[..., :3] + delta
    rgb = torch.max(torch.min(rgb, init[..., :3] + eps), init[..., :3] - eps)
    spatial_rgb[..., :3] = rgb
    return spatial_rgb
