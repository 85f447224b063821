"""ctypes binding of libnerfail_b200.so (the C ABI declared in include/nerfail_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, a RuntimeError is
raised.  Nothing here imports or executes anything under oracle/.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

_LIB_PATH = Path(__file__).resolve().parent / "_lib" / "libnerfail_b200.so"
_lib = None

c_f32p = C.c_void_p
c_i64 = C.c_int64
c_int = C.c_int
c_ptr = C.c_void_p

# name -> (restype, argtypes); must list every symbol of include/nerfail_b200.h (tests/test_abi.py checks)
SIGNATURES = {
    "nfb_abi_version": (c_int, []),
    "nfb_last_error": (C.c_char_p, []),
    "nfb_launch_count": (C.c_uint64, []),
    "nfb_device_cc": (c_int, []),
    "nfb_get_rays": (c_int, [c_int, c_int, c_ptr, c_ptr, C.c_float, C.c_float, c_ptr, c_ptr]),
    "nfb_rays_from_batch": (c_int, [c_ptr, c_ptr, c_i64, C.c_float, C.c_float, c_ptr, c_ptr]),
    "nfb_mse_loss2": (c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_ptr]),
    "nfb_coarse_z": (c_int, [c_ptr, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr]),
    "nfb_philox_uniform": (c_int, [C.c_uint64, C.c_uint64, C.c_uint32, c_i64, c_ptr, c_ptr]),
    "nfb_coarse_z_rng": (c_int, [c_ptr, c_int, c_int, c_int, C.c_uint64, C.c_uint64, c_ptr, c_ptr]),
    "nfb_mlp_create": (c_int, [C.POINTER(c_ptr), c_int, c_int, c_int, c_int, c_int]),
    "nfb_mlp_update": (c_int, [c_ptr, c_ptr, c_i64, c_ptr]),
    "nfb_mlp_destroy": (c_int, [c_ptr]),
    "nfb_mlp_param_count": (c_i64, [c_ptr]),
    "nfb_mlp_fwd": (c_int, [c_ptr, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_ptr, c_ptr]),
    "nfb_mlp_status": (c_int, [c_ptr]),
    "nfb_mlp_poll": (c_int, [c_ptr]),
    "nfb_mlp_debug_raise_abort": (c_int, [c_ptr]),
    "nfb_mlp_fwd_debug": (c_int, [c_ptr, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_ptr, c_int, c_ptr, c_ptr]),
    "nfb_mlp_train_tiles": (c_i64, [c_i64]),
    "nfb_mlp_fwd_train": (c_int, [c_ptr, c_ptr, c_ptr, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr]),
    "nfb_mlp_bwd_data": (c_int, [c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_ptr]),
    "nfb_mlp_fwd_trace": (c_int, [c_ptr, c_ptr, c_ptr, c_int, c_int, c_ptr, c_ptr, c_ptr]),
    "nfb_embed": (c_int, [c_ptr, c_i64, c_int, c_ptr, c_int, c_int, c_i64, c_ptr]),
    "nfb_linear_fwd": (c_int, [c_ptr, c_int, c_ptr, c_int, c_ptr, c_i64, c_int, c_int, c_int, c_ptr, c_int, c_ptr]),
    "nfb_linear_bwd_data": (c_int, [c_ptr, c_int, c_ptr, c_int, c_int, c_ptr, c_int, c_i64, c_int, c_int, c_ptr, c_int, c_int, c_ptr]),
    "nfb_linear_bwd_weight": (c_int, [c_ptr, c_int, c_ptr, c_int, c_int, c_ptr, c_int, c_i64, c_int, c_int, c_ptr, c_int, c_ptr, c_ptr, c_i64, c_ptr]),
    "nfb_linear_bwd_weight_workspace": (c_i64, [c_i64, c_int, c_int]),
    "nfb_mlp_bwd_weights": (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_ptr]),
    "nfb_mlp_bwd": (c_int, [c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "nfb_wgrad_bf16": (c_int, [c_ptr, c_i64, c_int, c_ptr, c_i64, c_int, c_i64, c_ptr, c_int, c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr]),
    "nfb_render_rays_workspace_bytes": (C.c_size_t, [c_int, c_int, c_int]),
    "nfb_render_rays_fwd": (c_int, [c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_int, c_ptr, c_ptr,
                                    c_int, C.c_uint64, C.c_uint64] + [c_ptr] * 8 + [c_ptr, C.c_size_t, c_ptr]),
    "nfb_render_rays_train_workspace_bytes": (C.c_size_t, [c_int, c_int, c_int]),
    "nfb_render_rays_train_raw_offset": (C.c_size_t, [c_int, c_int, c_int]),
    "nfb_render_rays_train_fwd": (c_int, [c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_int, c_ptr, c_ptr] + [c_ptr] * 7
                                  + [c_ptr, C.c_size_t, c_ptr]),
    "nfb_render_rays_bwd": (c_int, [c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_int] + [c_ptr] * 6 + [c_ptr, c_ptr]
                            + [c_ptr, C.c_size_t, c_ptr]),
    "nfb_attack_pack_rgb": (c_int, [c_ptr, c_ptr, c_i64, c_ptr, c_ptr]),
    "nfb_attack_sign_step": (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_i64, C.c_float, C.c_float, c_ptr]),
    "nfb_rgba_to_chw": (c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_i64, C.c_float, c_ptr, c_ptr]),
    "nfb_chw_to_rgba": (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_ptr]),
    "nfb_adam_step": (c_int, [c_ptr, c_int, c_i64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, c_ptr]),
    "nfb_adam_step_scalars": (c_int, [c_i64, C.c_double, C.c_double, C.c_double, c_ptr]),
    "nfb_adam_step_dev": (c_int, [c_ptr, c_int, c_ptr, C.c_double, C.c_double, C.c_double, C.c_double, c_ptr]),
    "nfb_composite_fwd": (c_int, [c_ptr, c_ptr, c_ptr, c_int, c_ptr, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "nfb_composite_bwd": (c_int, [c_ptr, c_ptr, c_ptr, c_int, c_ptr, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "nfb_sample_pdf": (c_int, [c_ptr, c_ptr, c_int, c_ptr, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr]),
    "nfb_hierarchical": (c_int, [c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr]),
    "nfb_hierarchical_rng": (c_int, [c_ptr, c_ptr, C.c_uint64, C.c_uint64, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr]),
    "nfb_knn8": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_ptr]),
    "nfb_knn_grid_cells": (c_i64, []),
    "nfb_knn_grid_build": (c_int, [c_ptr, c_i64, c_ptr, C.c_float, c_ptr, c_ptr, c_ptr, c_ptr]),
    "nfb_knn8_grid": (c_int, [c_ptr, c_i64, c_ptr, c_ptr, c_ptr, C.c_float, C.c_float, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "nfb_peer_alloc": (c_int, [C.c_size_t, C.POINTER(c_ptr)]),
    "nfb_peer_free": (c_int, [c_ptr]),
    "nfb_peer_export": (c_int, [c_ptr, c_ptr]),
    "nfb_peer_import": (c_int, [c_ptr, C.POINTER(c_ptr)]),
    "nfb_peer_close": (c_int, [c_ptr]),
    "nfb_peer_flag_bytes": (c_int, []),
    "nfb_peer_create": (c_int, [C.POINTER(c_ptr), c_int, c_int, c_ptr, c_ptr, c_ptr]),
    "nfb_peer_destroy": (c_int, [c_ptr]),
    "nfb_peer_status": (c_int, [c_ptr]),
    "nfb_attack_exchange_step": (c_int, [c_ptr, c_ptr, c_i64, C.c_float, C.c_float, c_ptr]),
    "nfb_adam_exchange_step": (c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_ptr, C.c_double, C.c_double, C.c_double, C.c_double, c_ptr]),
    "nfb_gauss_weights": (c_int, [c_ptr, c_i64, c_i64, C.c_float, c_ptr, c_ptr]),
    "nfb_gauss_gather_fwd": (c_int, [c_ptr, c_i64, c_ptr, c_ptr, c_i64, c_i64, C.c_float, c_ptr, c_ptr, c_ptr, c_ptr]),
    "nfb_gauss_scatter_bwd_batched": (c_int, [c_ptr, c_int, c_ptr, c_ptr, c_ptr, c_i64, c_i64, C.c_float, c_i64, c_ptr, c_ptr]),
    "nfb_resize_max_taps": (c_int, [c_int, c_int, c_int, c_int]),
    "nfb_resize_weights": (c_int, [c_int, c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr]),
    "nfb_rgba_to_chw_resized": (c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_int, c_int, c_int, c_int, C.c_float,
                                        c_ptr, c_ptr, c_ptr, c_int, c_ptr, c_ptr, c_ptr, c_int, c_ptr, c_ptr]),
    "nfb_chw_resized_to_rgba": (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_int, c_int, c_int, c_int,
                                        c_ptr, c_ptr, c_ptr, c_int, c_ptr, c_ptr, c_ptr, c_int, c_ptr, c_ptr]),
    "nfb_gauss_scatter_bwd": (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_i64, C.c_float, c_i64, c_ptr, c_ptr]),
}


def lib_path() -> Path:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load the shared library (once). Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise RuntimeError(
            f"{_LIB_PATH} is missing: build it with `python -m nerfail_b200.build` "
            "(nerfail_b200 has no CPU or PyTorch fallback)")
    lib = C.CDLL(str(_LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.nfb_abi_version() != 2:
        raise RuntimeError("libnerfail_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def last_error() -> str:
    return (load().nfb_last_error() or b"").decode()


def check(rc: int, what: str) -> None:
    if rc != 0:
        names = {-1: "NFB_E_ARG", -2: "NFB_E_UNSUPPORTED", -3: "NFB_E_CUDA"}
        raise RuntimeError(f"{what} failed with {names.get(rc, rc)}: {last_error()}")


def ptr(t: torch.Tensor | None) -> int | None:
    """Device pointer of a contiguous CUDA tensor (None passes NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("nerfail_b200 kernels need CUDA tensors (there is no CPU path)")
    if not t.is_contiguous():
        raise RuntimeError("nerfail_b200 kernels need contiguous tensors")
    return t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def launch_count() -> int:
    return int(load().nfb_launch_count())
