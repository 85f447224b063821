"""Host-side data feeding: the Blender loader against outputs of the unmodified reference loader (golden fixture
minted by tests/golden/make_golden_blender.py on the scene oracle/synth.write_blender_scene writes)."""
import os

import numpy as np
import torch

from conftest import golden
from oracle import synth


def test_load_blender_data_matches_reference_loader(tmp_path):
    from nerfail_b200 import data
    g = golden("blender.npz")
    root, att = str(tmp_path / "scene"), str(tmp_path / "attacked")
    synth.write_blender_scene(root, 8, 8, (3, 2, 4), seed=0, train_dir=att)
    cases = (("plain", {}), ("skip2", {"testskip": 2}), ("half", {"half_res": True}), ("attacked", {"train_dir": att}),
             ("attacked_half", {"train_dir": att, "half_res": True, "testskip": 0}))
    for tag, kw in cases:
        imgs, poses, render_poses, hwf, i_split = data.load_blender_data(root, **kw)
        if "train_dir" in kw:
            assert isinstance(imgs, list) and len(imgs) == 2
            assert np.array_equal(imgs[0], g[f"{tag}.train_imgs"]) and imgs[0].dtype == g[f"{tag}.train_imgs"].dtype
            imgs = imgs[1]
        assert np.array_equal(imgs, g[f"{tag}.imgs"]) and imgs.dtype == g[f"{tag}.imgs"].dtype, tag
        assert np.array_equal(poses, g[f"{tag}.poses"]) and poses.dtype == np.float32
        assert np.array_equal(render_poses.numpy(), g[f"{tag}.render_poses"])
        assert np.array_equal(np.asarray(hwf, np.float64), g[f"{tag}.hwf"])
        for k in range(3):
            assert np.array_equal(i_split[k], g[f"{tag}.split{k}"])
    assert np.array_equal(data.pose_spherical(33.0, -30.0, 4.0).numpy(), g["pose_spherical"])
    # the oracle's synthetic camera ring is the same construction
    assert np.allclose(synth.pose_spherical(33.0, -30.0, 4.0), g["pose_spherical"], atol=1e-6)


def test_white_background_composite():
    from nerfail_b200 import data
    rng = np.random.default_rng(0)
    imgs = rng.random((2, 4, 4, 4)).astype(np.float32)
    want = imgs[..., :3] * imgs[..., -1:] + (1. - imgs[..., -1:])            # run_nerf.py:585
    assert np.array_equal(data.white_background(imgs), want)
    pair = data.white_background([imgs[:1], imgs[1:]])                         # run_nerf.py:587-596 (train_dir pair)
    assert np.array_equal(pair[0], want[:1]) and np.array_equal(pair[1], want[1:])
