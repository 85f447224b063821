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
                st["step"] += 1
                steps.add(int(st["step"].item()))
                table[i] = _AdamTensor(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel())
                keep.append(g)
            if len(steps) != 1:
                raise RuntimeError("parameters of one group must share their step count")
            b1, b2 = group["betas"]
            with torch.cuda.device(live[0].device):
                check(lib.nfb_adam_step(C.cast(table, C.c_void_p), len(live), steps.pop(), float(group["lr"]), float(b1), float(b2),
                                        float(group["eps"]), self.grad_scale, stream()), "nfb_adam_step")
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
