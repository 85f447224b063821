"""Philox4x32-10 (Salmon, Moraes, Dror, Shaw 2011) in numpy: the checker for the kernel-side generator of csrc/sampling.cu."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint64(0x9E3779B9), np.uint64(0xBB67AE85)
MASK, S32 = np.uint64(0xFFFFFFFF), np.uint64(32)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over uint64 arrays holding 32-bit words; returns the four output words."""
    c0, c1, c2, c3 = (np.asarray(v, np.uint64) & MASK for v in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0) & MASK, np.uint64(k1) & MASK
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        h0, l0, h1, l1 = p0 >> S32, p0 & MASK, p1 >> S32, p1 & MASK
        c0, c1, c2, c3 = (h1 ^ c1 ^ k0) & MASK, l1, (h0 ^ c3 ^ k1) & MASK, l0
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0, c1, c2, c3


def uniform(seed: int, offset: int, stream_id: int, n: int) -> np.ndarray:
    """Element p = word p & 3 at counter (p >> 2, (p >> 34), stream_id ^ offset_hi, offset_lo), key = seed; 24-bit [0,1)."""
    p = np.arange(n, dtype=np.uint64)
    ctr = p >> np.uint64(2)
    words = philox4x32_10(ctr & MASK, ctr >> S32, np.full(n, (stream_id ^ (offset >> 32)) & 0xFFFFFFFF, np.uint64),
                          np.full(n, offset & 0xFFFFFFFF, np.uint64), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    sel = (p & np.uint64(3)).astype(np.int64)
    w = np.choose(sel, words)
    return ((w >> np.uint64(8)).astype(np.float64) * 2.0 ** -24).astype(np.float32)
