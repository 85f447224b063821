"""Builds libnerfail_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

The shared object is git-ignored but travels to the GPU box with the gpurun snapshot; nothing is
JIT-compiled at import time.  `python -m nerfail_b200.build` rebuilds from the command line.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OUT_DIR = PKG / "_lib"
LIB = OUT_DIR / "libnerfail_b200.so"
STAMP = OUT_DIR / "build.stamp"

SOURCES = ["api.cu", "composite.cu", "sampling.cu", "gauss.cu", "resize.cu", "peer.cu", "knn.cu", "linear.cu", "mlp_fused.cu", "wgrad.cu", "optim.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libnerfail_b200.so cannot be built")
    return exe


def _fingerprint() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.inl")) + [PKG.parent / "include" / "nerfail_b200.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    return LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == _fingerprint()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu for sm_100a and link the C-ABI shared library. Returns its path."""
    if not force and is_current():
        return LIB
    OUT_DIR.mkdir(exist_ok=True)
    nvcc = _nvcc()
    objs = []

    def compile_one(src: str) -> Path:
        obj = OUT_DIR / (src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *map(str, objs)]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    STAMP.write_text(_fingerprint())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
