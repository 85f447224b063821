"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol include/nerfail_b200.h
declares, the ctypes table covers them, and the product refuses to run without CUDA (no silent fallback)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import REPO


def header_symbols():
    text = open(os.path.join(REPO, "include", "nerfail_b200.h")).read()
    return sorted(set(re.findall(r"NFB_API[^;]*?\b(nfb_\w+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    from nerfail_b200 import build, _lib
    path = build.build()
    assert path.exists()
    lib = ctypes.CDLL(str(path))
    names = header_symbols()
    assert len(names) >= 26
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"
    assert _lib.load().nfb_abi_version() == 2


def test_every_declared_symbol_cites_the_reference_interface_it_replaces():
    """include/nerfail_b200.h: the comment that governs each NFB_API declaration names the reference file:line the entry
    point replaces, or says that it has no reference counterpart (plumbing, validation and profiling entries)."""
    text = re.sub(r"(?m)^#.*$", "", open(os.path.join(REPO, "include", "nerfail_b200.h")).read())   # drop #define NFB_API ...
    last, bad, n = "", [], 0
    for m in re.finditer(r"/\*(.*?)\*/|(NFB_API[^;]*;)", text, re.S):
        if m.group(1) is not None:
            last = m.group(1)
            continue
        n += 1
        name = re.search(r"\b(nfb_\w+)\s*\(", m.group(2)).group(1)
        if not (re.search(r"\w+\.py:\d+", last) or "no reference counterpart" in last):
            bad.append(name)
    assert n == len(header_symbols()) and not bad, bad


def test_error_reporting_without_compute():
    from nerfail_b200 import _lib
    lib = _lib.load()
    # argument validation happens before any CUDA call, so it is testable without a GPU
    rc = lib.nfb_composite_fwd(None, None, None, 3, None, 4, 8, 0, None, None, None, None, None, None, None)
    assert rc == -1 and b"composite_fwd" in lib.nfb_last_error()
    rc = lib.nfb_hierarchical(1, 1, None, 4, 2, 128, 1, None, None, None)
    assert rc == -2 and b"Sc" in lib.nfb_last_error()
    rc = lib.nfb_knn8(1, 10, 1, 3, 1, None, None, None)
    assert rc == -1
    h = ctypes.c_void_p()
    rc = lib.nfb_mlp_create(ctypes.byref(h), 4, 128, 63, 27, 4)
    assert rc == -2 and b"D=8" in lib.nfb_last_error()


def test_no_cpu_fallback():
    import nerfail_b200 as nb
    from nerfail_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.composite_fwd(torch.zeros(2, 4, 4), torch.zeros(2, 4), torch.zeros(2, 3))
    with pytest.raises(RuntimeError, match="CUDA"):
        nb.NeRF(use_viewdirs=True, input_ch=63, input_ch_views=27)(torch.zeros(3, 90))
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.knn8(torch.zeros(4, 3), torch.zeros(16, 3))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from nerfail_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "_LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(REPO, "nerfail_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"


def test_state_dict_keys_and_param_order_match_reference_layout():
    import nerfail_b200 as nb
    from oracle import synth
    net = nb.NeRF(D=8, W=256, input_ch=63, output_ch=5, skips=[4], input_ch_views=27, use_viewdirs=True)
    sd = synth.random_state_dict(0)
    assert list(net.state_dict().keys()) == list(sd.keys())
    net.load_state_dict(sd)
    assert torch.equal(net.flat_params(), synth.flat_params(sd))
    assert net.flat_params().numel() == 595844
    assert _count() == 595844


def _count():
    from nerfail_b200 import _lib
    return int(_lib.load().nfb_mlp_param_count(None))


def test_philox_reference_known_answers():
    """The numpy Philox4x32-10 that checks the kernel-side generator (tests/philox_ref.py) against the known-answer vectors
    of the Random123 distribution (kat_vectors: philox4x32 10 rounds)."""
    import numpy as np
    from philox_ref import philox4x32_10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox4x32_10(*[np.array([c], np.uint64) for c in ctr], *key)
        assert tuple(int(g[0]) for g in got) == want
