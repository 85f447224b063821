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
                      minimise: bool = True, active_idx: Optional[torch.Tensor] = None,
                      exchange: Optional["PeerExchange"] = None) -> torch.Tensor:
    """The I-FGSM update of attack_NeRFail_S.py:357-392 applied AFTER the gradient all-reduce: sign step on the RGB
    channels where the point is active (A > 0), then clamp to init +- eps.  Identical on every rank by construction.
    With active_idx (= active_rows(spatial_rgb)) only those rows are exchanged and updated; the result is the same.
    With exchange (a PeerExchange whose `value` IS spatial_rgb and whose `grad` IS grad) the reduction, the update and the
    broadcast are one kernel per GPU over NVLink peer memory instead of an NCCL all-reduce plus an update kernel."""
    if exchange is not None:
        if spatial_rgb.data_ptr() != exchange.value.data_ptr() or grad.data_ptr() != exchange.grad.data_ptr():
            raise RuntimeError("attack_sign_step_(exchange=...): spatial_rgb / grad must be the exchange's value / grad buffers")
        exchange.attack_step(init.reshape(-1, 4).contiguous(), spatial_rgb.numel() // 4, step, eps, minimise)
        return spatial_rgb
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


# ------------------------------------------------------------------------------------------------------------------
# exchange steps over NVLink peer memory (csrc/peer.cu): reduce-scatter + update + all-gather in ONE kernel per GPU
# ------------------------------------------------------------------------------------------------------------------
class _DeviceMemory:
    """Raw device memory as a torch tensor without a copy (torch.as_tensor consumes __cuda_array_interface__)."""

    def __init__(self, ptr: int, n_floats: int, typestr: str = "<f4"):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": typestr, "data": (ptr, False), "version": 2}


class PeerExchange:
    """Symmetric (gradient, value, flags) buffers of the ranks of ONE node and the fused exchange kernels on them.

    Every rank allocates its buffers with nfb_peer_alloc, publishes CUDA IPC handles through the process group
    (all_gather_object — gloo or NCCL, host plumbing only) and opens the peers' buffers; afterwards no NCCL call is on the
    data path: `attack_step` / `adam_step` are one kernel per GPU that reads the peers' gradients and writes the peers'
    values through NVLink (csrc/peer.cu).  `grad` and `value` are this rank's buffers as flat fp32 tensors: accumulate the
    local gradient into `grad`, keep the replicated quantity (perturbation table / flat parameters) in `value`.
    With one rank the same kernels run without any peer traffic, so single-GPU code takes the same path."""

    def __init__(self, n_grad: int, n_value: int, device=None, group=None):
        import ctypes as C
        from . import _lib
        self._lib, self._C = _lib, C
        lib = _lib.load()
        self.device = torch.device(device if device is not None else ("cuda", torch.cuda.current_device()))
        self.rank, self.world = world() if group is None else (dist.get_rank(group), dist.get_world_size(group))
        if self.world > 8:
            raise RuntimeError("PeerExchange spans the GPUs of one node (at most 8 ranks)")
        pad4 = lambda n: (int(n) + 3) // 4 * 4
        self.n_grad, self.n_value = pad4(n_grad), pad4(n_value)
        # ONE allocation per rank, a multiple of 2 MiB (an IPC handle exports the whole underlying allocation, and small
        # cudaMalloc blocks share one): [flag words | gradient | value], each part 256-byte aligned
        al = lambda n: (n + 255) // 256 * 256
        off_grad = al(int(lib.nfb_peer_flag_bytes()))
        off_value = off_grad + al(self.n_grad * 4)
        total = (off_value + al(self.n_value * 4) + (1 << 21) - 1) >> 21 << 21
        self._opened = []
        with torch.cuda.device(self.device):
            p = C.c_void_p()
            _lib.check(lib.nfb_peer_alloc(total, C.byref(p)), "nfb_peer_alloc")
            self._local = p.value
            hbuf = C.create_string_buffer(64)
            _lib.check(lib.nfb_peer_export(p, hbuf), "nfb_peer_export")
            everyone = [None] * self.world
            if self.world > 1:
                dist.all_gather_object(everyone, (self.rank, hbuf.raw), group=group)
            else:
                everyone = [(0, hbuf.raw)]
            bases = [None] * self.world
            for r, hraw in everyone:
                if r == self.rank:
                    bases[r] = self._local
                else:
                    q = C.c_void_p()
                    _lib.check(lib.nfb_peer_import(C.create_string_buffer(hraw, 64), C.byref(q)), "nfb_peer_import")
                    self._opened.append(q.value)
                    bases[r] = q.value
            arrs = [(C.c_void_p * self.world)(*[b + o for b in bases]) for o in (off_grad, off_value, 0)]
            h = C.c_void_p()
            _lib.check(lib.nfb_peer_create(C.byref(h), self.rank, self.world, arrs[0], arrs[1], arrs[2]), "nfb_peer_create")
            self._h = h
            self.grad = torch.as_tensor(_DeviceMemory(self._local + off_grad, self.n_grad), device=self.device)
            self.value = torch.as_tensor(_DeviceMemory(self._local + off_value, self.n_value), device=self.device)
            if self.world > 1:
                # sanity check of the mapping before any kernel trusts it: every rank marks its own buffer, reads the
                # peers' marks through the opened pointers, and clears its mark again
                self.value[:4] = float(1000 + self.rank)
                torch.cuda.synchronize(self.device)
                dist.barrier(group=group)
                for r in range(self.world):
                    if r != self.rank:
                        seen = torch.as_tensor(_DeviceMemory(bases[r] + off_value, 4), device=self.device).cpu()
                        if not bool((seen == float(1000 + r)).all()):
                            raise RuntimeError(f"PeerExchange: rank {self.rank} does not see rank {r}'s buffer through its IPC mapping "
                                               f"(read {seen.tolist()})")
                dist.barrier(group=group)
                self.value[:4] = 0.0
                torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(group=group)           # every rank has opened every buffer before anyone launches

    def status(self) -> None:
        """Raises if a flag wait of an exchange kernel that has finished timed out (no synchronisation)."""
        self._lib.check(self._lib.load().nfb_peer_status(self._h), "nfb_peer_status")

    def attack_step(self, init: torch.Tensor, n_rows: int, step: float, eps: float, minimise: bool = True) -> None:
        """attack_NeRFail_S.py:348-392 for this iteration: `grad` [n_rows,4] summed over ranks, sign step on the rows of
        `value` [n_rows,4] with A > 0, clamped to init +- eps; every rank's `value` is updated, `grad` may be zeroed after."""
        assert n_rows * 4 <= self.n_grad and n_rows * 4 <= self.n_value and init.is_contiguous() and init.numel() == n_rows * 4
        with torch.cuda.device(self.device):
            self._lib.check(self._lib.load().nfb_attack_exchange_step(self._h, self._lib.ptr(init), n_rows,
                                                                      float(step if minimise else -step), float(eps),
                                                                      self._lib.stream()), "nfb_attack_exchange_step")

    def adam_step(self, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, n: int, step_scalars: torch.Tensor, betas, eps: float,
                  grad_scale: float) -> None:
        """run_nerf.py:791-792 for this step: `grad` [n] averaged over ranks (grad_scale = 1 / world), Adam on this rank's
        slice (the optimiser state is sharded), the new parameters written into every rank's `value` [n]."""
        with torch.cuda.device(self.device):
            self._lib.check(self._lib.load().nfb_adam_exchange_step(self._h, self._lib.ptr(exp_avg), self._lib.ptr(exp_avg_sq), n,
                                                                    self._lib.ptr(step_scalars), float(betas[0]), float(betas[1]),
                                                                    float(eps), float(grad_scale), self._lib.stream()),
                            "nfb_adam_exchange_step")

    def close(self) -> None:
        lib = self._lib.load()
        if getattr(self, "_h", None):
            torch.cuda.synchronize(self.device)
            if self.world > 1 and dist.is_initialized():
                dist.barrier()                  # nobody unmaps memory a peer's kernel may still touch
            lib.nfb_peer_destroy(self._h)
            self._h = None
            for p in self._opened:
                lib.nfb_peer_close(p)
            self.grad = self.value = None
            if self.world > 1 and dist.is_initialized():
                dist.barrier()                  # ... and nobody frees memory a peer still has mapped
            if self._local:
                lib.nfb_peer_free(self._local)
            self._opened, self._local = [], None

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self.world == 1:
                self.close()
        except Exception:
            pass


def shard_pixels(n_views: int, pixels_per_view: int, rank: int, world_size: int, quantum: int = 1):
    """Balanced split of a batch of views by PIXELS rather than whole views (100 views over 8 ranks: 12.5 views each
    instead of 13 / 12): [(view, pixel_begin, pixel_end)] of this rank, contiguous in (view, pixel) order.  The GaussNet
    kernels are per pixel, so a rank can take part of a view (its rows); `quantum` keeps cuts on multiples of e.g. a row."""
    total = n_views * pixels_per_view
    units = total // quantum
    b, e = shard_range(units, rank, world_size)
    b, e = b * quantum, (e * quantum if rank < world_size - 1 else total)
    out = []
    v = b // pixels_per_view
    while b < e:
        end = min(e, (v + 1) * pixels_per_view)
        out.append((v, b - v * pixels_per_view, end - v * pixels_per_view))
        b, v = end, v + 1
    return out


class PeerAdam:
    """Data-parallel optimiser step of NeRF retraining as ONE kernel per GPU (csrc/peer.cu: adam_exchange_kernel).

    Replaces `all_reduce(grads) ; optimizer.step()` (a data-parallel run_nerf.py:791-792): the parameters of the networks
    are re-pointed into a flat buffer in NVLink peer memory (state_dict order, coarse network first), the fused training
    backward accumulates its flat gradient straight into the matching peer buffer (`net._grad_sink`), and `step()` lets
    every GPU average ITS 1/G slice of the gradients over the peers, run Adam on it (exp_avg / exp_avg_sq are sharded: a
    rank only ever touches its slice) and write the new parameters into every rank's copy.  No NCCL call, no replicated
    full-size Adam.  `optimizer` (nerfail_b200.optim.Adam over the same parameters) keeps providing lr, betas, the step
    count and the device-resident step scalars (graph mode), so learning-rate decay and CUDA-graph replay work as before;
    `state_for_checkpoint()` gathers the sharded moments back into torch.optim.Adam's state_dict layout."""

    def __init__(self, nets, optimizer, device=None, group=None):
        self.nets = [n for n in nets if n is not None]
        self.optimizer = optimizer
        params = [p for n in self.nets for p in n.ordered_params()]
        opt_params = [p for g in optimizer.param_groups for p in g["params"]]
        if len(opt_params) != len(params) or any(a is not b for a, b in zip(params, opt_params)):
            raise RuntimeError("PeerAdam: the optimizer must hold exactly the networks' parameters in state_dict order")
        self.device = torch.device(device if device is not None else params[0].device)
        self.n = sum(p.numel() for p in params)
        self.ex = PeerExchange(self.n, self.n, self.device, group)
        self.world = self.ex.world
        self.m = torch.zeros(self.ex.n_value, dtype=torch.float32, device=self.device)
        self.v = torch.zeros(self.ex.n_value, dtype=torch.float32, device=self.device)
        off = 0
        with torch.no_grad():
            for net in self.nets:
                begin = off
                for p in net.ordered_params():
                    k = p.numel()
                    self.ex.value[off:off + k].copy_(p.data.reshape(-1))
                    st = optimizer.state.get(p, {})
                    if "exp_avg" in st:                      # resume: take over the moments the optimizer already holds
                        self.m[off:off + k].copy_(st["exp_avg"].reshape(-1))
                        self.v[off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
                    p.data = self.ex.value[off:off + k].view(p.shape)
                    off += k
                net._grad_sink = self.ex.grad[begin:off]
                net.invalidate_fused()
            root = self.ex.grad[:off]                       # the sinks of all networks, contiguous: zeroed with one fill
            for net in self.nets:
                net._grad_sink_root = root
        self.params = params
        optimizer.enable_graph_mode(self.device)

    def step(self, begin: bool = True) -> None:
        """After loss.backward(): average, Adam, broadcast.  begin=False when the caller (GraphedTrainStep) has already
        advanced the optimizer's step count and uploaded this step's scalars (optimizer.begin_step())."""
        if begin:
            self.optimizer.begin_step()
        g = self.optimizer.param_groups[0]
        self.ex.adam_step(self.m, self.v, self.n, self.optimizer._dyn[0], g["betas"], g["eps"], 1.0 / self.world)
        for p in self.params:
            torch.autograd.graph.increment_version(p)

    def state_for_checkpoint(self) -> dict:
        """torch.optim.Adam state_dict with the full moments (each rank contributes its slice; collective)."""
        m, v = self.m.clone(), self.v.clone()
        if self.world > 1:
            n4 = (self.n + 3) // 4
            per = (n4 + self.world - 1) // self.world * 4
            own = torch.zeros_like(m)
            b, e = per * self.ex.rank, min(self.ex.n_value, per * (self.ex.rank + 1))
            own[b:e] = m[b:e]
            dist.all_reduce(own); m = own
            own = torch.zeros_like(v); own[b:e] = v[b:e]
            dist.all_reduce(own); v = own
        off = 0
        for p in self.params:
            st = self.optimizer.state[p]
            k = p.numel()
            st["exp_avg"], st["exp_avg_sq"] = m[off:off + k].view(p.shape).clone(), v[off:off + k].view(p.shape).clone()
            off += k
        return self.optimizer.state_dict()

    def close(self) -> None:
        """Gives the parameters their own storage back and releases the peer memory."""
        with torch.no_grad():
            for p in self.params:
                p.data = p.data.clone()
        for n in self.nets:
            n._grad_sink = None
            n._grad_sink_root = None
            n.invalidate_fused()
        self.ex.close()
