"""Stages the UNMODIFIED reference render path under oracle/_ref/ so that it can run on the GPU box's host cores.

    python oracle/make_ref.py            (build container only: copies from /root/reference)

The reference is a Python program (no build step): the files of its render path are packed byte for byte from where they
lie under /root/reference into ONE archive, oracle/_ref/nerf_pytorch_ref.zip (with their SHA-256 sums next to it) — the
"built artefact" of a Python reference.  oracle/_ref/ is git-ignored (the sources never enter this repository's history)
and NOT gpurun-ignored, so the archive travels to the GPU box like a built .so would; Python imports straight from it
(zipimport).  bench.py's `--impl reference`
arm and its `cpu_baseline` leg import run_nerf / run_nerf_helpers from there (kind "reference"); when the directory is
missing they fall back to the oracle port (kind "port").  Nothing else may import it: it is the checker / the baseline,
never the product.  `ref_modules()` registers the empty stand-ins for imageio / matplotlib / configargparse that the
reference imports at module scope without using them in any arithmetic, then imports the two modules.
"""
from __future__ import annotations

import hashlib
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
ARCHIVE = os.path.join(DST, "nerf_pytorch_ref.zip")
SRC = os.path.join(os.environ.get("NERFAIL_REFERENCE", "/root/reference"), "Create_spatial_point_set", "nerf_pytorch")
FILES = ["run_nerf.py", "run_nerf_helpers.py", "load_blender.py", "load_llff.py", "load_deepvoxels.py", "load_LINEMOD.py", "LICENSE"]


def stage() -> str | None:
    """Copies the files; returns the destination, or None when the reference tree is not present (GPU box)."""
    if not os.path.isdir(SRC):
        return None
    import zipfile
    os.makedirs(DST, exist_ok=True)
    for f in os.listdir(DST):                    # a previous staging may have left loose files
        if f.endswith(".py") or f == "LICENSE":
            os.remove(os.path.join(DST, f))
    lines = []
    with zipfile.ZipFile(ARCHIVE, "w", zipfile.ZIP_DEFLATED) as z:
        for f in FILES:
            data = open(os.path.join(SRC, f), "rb").read()
            z.writestr(zipfile.ZipInfo(f, date_time=(2020, 1, 1, 0, 0, 0)), data)       # fixed timestamps: reproducible archive
            lines.append(f"{hashlib.sha256(data).hexdigest()}  {f}")
    with open(os.path.join(DST, "SHA256SUMS"), "w") as fp:
        fp.write("\n".join(lines) + "\n")
    return DST


def available() -> bool:
    return os.path.isfile(ARCHIVE)


def ref_modules():
    """(run_nerf, run_nerf_helpers) of the staged, unmodified reference."""
    for name in ("imageio", "matplotlib", "matplotlib.pyplot", "configargparse"):
        try:
            importlib.import_module(name)
        except Exception:
            m = types.ModuleType(name)
            sys.modules[name] = m
            if "." in name:
                setattr(sys.modules[name.split(".")[0]], name.split(".")[1], m)
    if ARCHIVE not in sys.path:
        sys.path.insert(0, ARCHIVE)
    import run_nerf
    import run_nerf_helpers
    return run_nerf, run_nerf_helpers


if __name__ == "__main__":
    d = stage()
    print("staged the reference render path under", d) if d else print("no reference tree at", SRC)
