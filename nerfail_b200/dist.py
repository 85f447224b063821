"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the GPU box, gloo in CPU tests).

The reference is single-process; what shards and what must be exchanged follows SURVEY.md §8e:
  * render sweeps (run_nerf.py:151-154): views are independent -> shard by view, no collective;
  * one view: rays are independent -> contiguous ray ranges, no collective;
  * NeRFail-S attack iteration (attack_NeRFail_S.py:304-392): perturbation replicated, views sharded, ONE all-reduce
    (sum) of grad_spatial_rgb [P,H,W,4] before the sign step, so every rank applies the identical update;
  * NeRF retraining (run_nerf.py:776-792): rays sharded, ONE all-reduce of the parameter gradients per network.
"""
from __future__ import annotations

from typing import Optional, Iterable, List, Sequence, Tuple

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


def broadcast_params_(modules, src: int = 0) -> None:
    """Initial parameter sync of data-parallel retraining: every rank takes rank `src`'s weights.  The broadcast writes
    through `p.data`, which does not bump the version counters NeRF.fused() watches, so the packed bf16 images are
    invalidated explicitly (NeRF.invalidate_fused)."""
    if world()[1] > 1:
        for m in modules:
            for p in m.parameters():
                dist.broadcast(p.data, src)
    for m in modules:
        if hasattr(m, "invalidate_fused"):
            m.invalidate_fused()


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


def active_rows(spatial_rgb: torch.Tensor) -> torch.Tensor:
    """Indices of the perturbation-table rows the sign step can change: A > 0 (attack_NeRFail_S.py:357-372 multiplies the
    step by that mask, and A itself is never updated, so the set is fixed for a whole attack)."""
    return torch.nonzero(spatial_rgb.reshape(-1, 4)[:, 3] > 0).reshape(-1)


def allreduce_active_rgb(grad: torch.Tensor, active_idx: torch.Tensor) -> torch.Tensor:
    """The part of grad_spatial_rgb the update consumes — the RGB columns of the active rows — packed to [n_active, 3] and
    summed over ranks: 0.75 x (active fraction) of the 30.72 MB a full all-reduce would move (0.3 x for the ~40 % of
    pixels an object covers).  For links slower than NVLink: on an 8 x B200 NVSwitch node the exchange is latency-bound and
    packing buys nothing (0.93 ms per 100-view iteration against 0.87-1.00 ms with the full table), so bench.py keeps the table."""
    if grad.is_cuda:
        from . import _lib
        g4 = grad.reshape(-1, 4)
        packed = torch.empty((active_idx.numel(), 3), dtype=torch.float32, device=grad.device)
        with torch.cuda.device(grad.device):
            _lib.check(_lib.load().nfb_attack_pack_rgb(_lib.ptr(g4), _lib.ptr(active_idx), active_idx.numel(), _lib.ptr(packed),
                                                       _lib.stream()), "nfb_attack_pack_rgb")
    else:                                          # host tensors: the gloo tests of the exchange logic
        packed = grad.reshape(-1, 4).index_select(0, active_idx)[:, :3].contiguous()
    allreduce_sum_(packed)
    return packed


def attack_sign_step_(spatial_rgb: torch.Tensor, grad: torch.Tensor, init: torch.Tensor, step: float, eps: float,
                      minimise: bool = True, active_idx: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The I-FGSM update of attack_NeRFail_S.py:357-392 applied AFTER the gradient all-reduce: sign step on the RGB
    channels where the point is active (A > 0), then clamp to init +- eps.  Identical on every rank by construction.
    With active_idx (= active_rows(spatial_rgb)) only those rows are exchanged and updated; the result is the same."""
    if active_idx is not None:
        g = allreduce_active_rgb(grad, active_idx)
        rows, rows0 = spatial_rgb.reshape(-1, 4), init.reshape(-1, 4)
        if rows.data_ptr() != spatial_rgb.data_ptr():
            raise RuntimeError("attack_sign_step_: spatial_rgb must be contiguous for the in-place row update")
        if rows.is_cuda:
            from . import _lib
            with torch.cuda.device(rows.device):
                _lib.check(_lib.load().nfb_attack_sign_step(_lib.ptr(rows), _lib.ptr(rows0.contiguous()), _lib.ptr(active_idx), _lib.ptr(g),
                                                            active_idx.numel(), float(step if minimise else -step), float(eps),
                                                            _lib.stream()), "nfb_attack_sign_step")
            return spatial_rgb
        rgb = rows[active_idx, :3]
        rgb0 = rows0[active_idx, :3]
        rgb = rgb - step * torch.sign(g) if minimise else rgb + step * torch.sign(g)
        rows[active_idx, :3] = torch.max(torch.min(rgb, rgb0 + eps), rgb0 - eps)
        return spatial_rgb
    allreduce_sum_(grad)
    active = (spatial_rgb[..., 3:4] > 0).to(spatial_rgb.dtype)
    delta = step * torch.sign(grad[..., :3]) * active
    rgb = spatial_rgb[..., :3] - delta if minimise else spatial_rgb[..., :3] + delta
    rgb = torch.max(torch.min(rgb, init[..., :3] + eps), init[..., :3] - eps)
    spatial_rgb[..., :3] = rgb
    return spatial_rgb
