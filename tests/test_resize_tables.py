"""nfb_resize_weights (a HOST function of the C ABI: the per-axis tap tables of the fused Resize kernel) against
torchvision.transforms.Resize on the CPU — the op GaussNet.py:147-154 applies to the classifier input.  No GPU needed:
the tables are expanded to dense matrices and applied with two matmuls."""
import ctypes as C

import numpy as np
import pytest
import torch

from nerfail_b200 import _lib


def dense(in_size, out_size, antialias, transposed):
    lib = _lib.load()
    maxk = lib.nfb_resize_max_taps(in_size, out_size, int(antialias), int(transposed))
    n = in_size if transposed else out_size
    start, count, w = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros((n, maxk), np.float32)
    assert lib.nfb_resize_weights(in_size, out_size, int(antialias), int(transposed), maxk, start.ctypes.data, count.ctypes.data,
                                  w.ctypes.data) == 0, _lib.last_error()
    m = np.zeros((out_size, in_size), np.float64)
    for i in range(n):
        for j in range(count[i]):
            if transposed:
                m[start[i] + j, i] = w[i, j]
            else:
                m[i, start[i] + j] = w[i, j]
    return m, int(count.max())


@pytest.mark.parametrize("in_size,out_size", [(800, 299), (800, 224), (100, 299), (64, 64), (37, 11), (5, 1)])
@pytest.mark.parametrize("antialias", [False, True])
def test_tables_reproduce_torchvision_resize(in_size, out_size, antialias):
    from torchvision.transforms import Resize
    g = torch.Generator().manual_seed(in_size * 1000 + out_size)
    img = torch.rand(2, 3, in_size, in_size, generator=g) * 255
    ref = Resize([out_size, out_size], antialias=antialias)(img).numpy().astype(np.float64)
    m, taps = dense(in_size, out_size, antialias, False)
    got = m @ img.numpy().astype(np.float64) @ m.T                                 # [out,in] @ [B,3,in,in] @ [in,out]
    assert np.abs(got - ref).max() <= 2e-4 * 255, np.abs(got - ref).max()          # fp32 weights / accumulation order
    np.testing.assert_allclose(m.sum(1), 1.0, atol=1e-5)                           # every output is a convex combination
    mt, _ = dense(in_size, out_size, antialias, True)                              # the adjoint's tables are the same matrix
    assert np.array_equal(m, mt)


def test_max_taps_bound_is_respected_and_small_maxk_is_refused():
    lib = _lib.load()
    k = lib.nfb_resize_max_taps(800, 299, 1, 0)
    _, used = dense(800, 299, True, False)
    assert used <= k <= used + 2
    start, count, w = np.zeros(299, np.int32), np.zeros(299, np.int32), np.zeros((299, 2), np.float32)
    assert lib.nfb_resize_weights(800, 299, 1, 0, 2, start.ctypes.data, count.ctypes.data, w.ctypes.data) == -1
