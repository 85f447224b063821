"""Generates tests/golden/blender.npz: the UNMODIFIED reference loader (load_blender.py) run in the build container on
the tiny scene oracle/synth.write_blender_scene writes.  imageio is not installed here; the stub's imread returns what
imageio returns for an RGBA PNG (uint8 [H,W,4], RGBA order), decoded by cv2.
    python tests/golden/make_golden_blender.py"""
import os
import sys
import tempfile
import types

import cv2
import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("NERFAIL_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)
from oracle import synth  # noqa: E402

stub = types.ModuleType("imageio")
stub.imread = lambda f: cv2.imread(f, cv2.IMREAD_UNCHANGED)[..., [2, 1, 0, 3]]
sys.modules["imageio"] = stub
sys.path.insert(0, os.path.join(REF, "Create_spatial_point_set", "nerf_pytorch"))
import load_blender  # noqa: E402

out = {}
with tempfile.TemporaryDirectory() as d:
    root, att = os.path.join(d, "scene"), os.path.join(d, "attacked")
    synth.write_blender_scene(root, 8, 8, (3, 2, 4), seed=0, train_dir=att)
    for tag, kw in (("plain", {}), ("skip2", {"testskip": 2}), ("half", {"half_res": True}),
                    ("attacked", {"train_dir": att}), ("attacked_half", {"train_dir": att, "half_res": True, "testskip": 0})):
        imgs, poses, render_poses, hwf, i_split = load_blender.load_blender_data(root, **kw)
        if isinstance(imgs, list):
            out[f"{tag}.train_imgs"], out[f"{tag}.imgs"] = np.asarray(imgs[0]), np.asarray(imgs[1])
        else:
            out[f"{tag}.imgs"] = np.asarray(imgs)
        out[f"{tag}.poses"] = poses
        out[f"{tag}.render_poses"] = render_poses.numpy()
        out[f"{tag}.hwf"] = np.asarray(hwf, np.float64)
        for k, s in enumerate(i_split):
            out[f"{tag}.split{k}"] = s
out["pose_spherical"] = load_blender.pose_spherical(33.0, -30.0, 4.0).numpy()
np.savez_compressed(os.path.join(REPO, "tests", "golden", "blender.npz"), **out)
print("wrote blender.npz:", {k: v.shape for k, v in out.items() if k.startswith("attacked_half")})
