"""Fused Adam for the NeRF retraining loop (run_nerf.py:213, :776-800).

`Adam` is a torch.optim.Optimizer whose state_dict has exactly torch.optim.Adam's layout (per-parameter `step`,
`exp_avg`, `exp_avg_sq`; the same param_group keys), so `optimizer.load_state_dict(ckpt['optimizer_state_dict'])`
(run_nerf.py:228) resumes reference checkpoints and its own checkpoints load into torch.optim.Adam.  The update of
every parameter tensor of both networks is ONE launch of nfb_adam_step.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .ops import check, stream


class _AdamTensor(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("numel", C.c_int64)]


class Adam(torch.optim.Optimizer):
    """Drop-in for torch.optim.Adam(params, lr, betas=(0.9, 0.999), eps=1e-8) as the reference constructs it.
    Unsupported torch options (weight_decay, amsgrad, maximize) raise instead of being ignored."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, *, maximize=False,
                 grad_scale=1.0):
        if weight_decay != 0 or amsgrad or maximize:
            raise NotImplementedError("nerfail_b200.optim.Adam implements the reference's configuration only "
                                      "(weight_decay=0, amsgrad=False, maximize=False)")
        if not 0.0 <= lr or not 0.0 <= eps or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("invalid Adam hyper-parameters")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, maximize=False, foreach=None,
                        capturable=False, differentiable=False, fused=None, decoupled_weight_decay=False)
        super().__init__(params, defaults)
        self.grad_scale = float(grad_scale)       # e.g. 1/world_size when gradients were all-reduced with SUM
        self._dyn = None                          # graph mode: (device [2] float32, pinned host mirror)

    # ---- CUDA-graph mode: step() launches the device-scalar kernel; begin_step() advances the step on the host ----
    def enable_graph_mode(self, device) -> None:
        """After this, step() reads lr / (1 - beta1^t) and 1 / sqrt(1 - beta2^t) from device memory (nfb_adam_step_dev), so
        a step captured in a CUDA graph can be replayed: call begin_step() before every step() or replay."""
        if self._dyn is None:
            # device scalars + a ring of pinned staging slots (a slot is rewritten only after its upload has finished)
            self._dyn = (torch.zeros(2, dtype=torch.float32, device=device), torch.zeros(16, 2, dtype=torch.float32).pin_memory())
            self._dyn_events, self._dyn_i = [None] * 16, 0

    def begin_step(self) -> int:
        """Graph mode: advance every parameter's step count (the torch.optim.Adam state_dict field) and upload the constants of
        this step for the current learning rate, stream-ordered before the kernel that reads them.  Returns the step."""
        if self._dyn is None:
            raise RuntimeError("begin_step() is for graph mode: call enable_graph_mode(device) first")
        lib = _lib.load()
        step = None
        for group in self.param_groups:
            for p in group["params"]:
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                step = int(st["step"].item()) if step is None else step
        group = self.param_groups[0]
        if len(self.param_groups) != 1:
            raise RuntimeError("graph mode supports one parameter group (the reference's optimizer has one)")
        dev, ring = self._dyn
        i = self._dyn_i
        self._dyn_i = (i + 1) % ring.shape[0]
        if self._dyn_events[i] is not None:
            self._dyn_events[i].synchronize()
        host = ring[i]
        check(lib.nfb_adam_step_scalars(step, float(group["lr"]), float(group["betas"][0]), float(group["betas"][1]),
                                        host.data_ptr()), "nfb_adam_step_scalars")
        dev.copy_(host, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev.device))
        self._dyn_events[i] = ev
        return step

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for group in self.param_groups:
            live = [p for p in group["params"] if p.grad is not None]
            if not live:
                continue
            steps = set()
            table = (_AdamTensor * len(live))()
            keep = []                                  # contiguous gradient copies must outlive the launch call
            for i, p in enumerate(live):
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("nerfail_b200.optim.Adam needs contiguous fp32 CUDA parameters")
                g = p.grad
                if g.is_sparse or g.dtype != torch.float32:
                    raise RuntimeError("nerfail_b200.optim.Adam needs dense fp32 gradients")
                if not g.is_contiguous():
                    g = g.contiguous()
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)            # host scalar tensor, like torch (capturable=False)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                if self._dyn is None:
                    st["step"] += 1
                steps.add(int(st["step"].item()))
                table[i] = _AdamTensor(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel())
                keep.append(g)
            if len(steps) != 1:
                raise RuntimeError("parameters of one group must share their step count")
            b1, b2 = group["betas"]
            with torch.cuda.device(live[0].device):
                if self._dyn is not None:
                    check(lib.nfb_adam_step_dev(C.cast(table, C.c_void_p), len(live), self._dyn[0].data_ptr(), float(b1), float(b2),
                                                float(group["eps"]), self.grad_scale, stream()), "nfb_adam_step_dev")
                else:
                    check(lib.nfb_adam_step(C.cast(table, C.c_void_p), len(live), steps.pop(), float(group["lr"]), float(b1),
                                            float(b2), float(group["eps"]), self.grad_scale, stream()), "nfb_adam_step")
            for p in live:
                torch.autograd.graph.increment_version(p)     # the kernel wrote through raw pointers: repack triggers on _version
        return loss


def decayed_lrate(lrate: float, lrate_decay: int, global_step: int, decay_rate: float = 0.1) -> float:
    """run_nerf.py:796-798: lrate * decay_rate ** (global_step / (lrate_decay * 1000))."""
    return lrate * (decay_rate ** (global_step / (lrate_decay * 1000)))


def set_lrate(optimizer: torch.optim.Optimizer, new_lrate: float) -> None:
    """run_nerf.py:799-800."""
    for group in optimizer.param_groups:
        group["lr"] = new_lrate
