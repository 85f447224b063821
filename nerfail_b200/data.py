"""Blender-scene loader with the reference's return contract, PNG decoding on a thread pool.

Reference: Create_spatial_point_set/nerf_pytorch/load_blender.py:29-110 (`pose_spherical`, `load_blender_data` with
NeRFail's `train_dir` override :37,62-63,69-70,80-81,107-108) and the white-background composite of run_nerf.py:584-596.
Host-side data feeding: no arithmetic of the render path lives here.  The reference decodes 400
800x800 PNGs one after the other with imageio; cv2.imread releases the GIL, so the same files decode on all host cores.
"""
from __future__ import annotations

import json
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch


def pose_spherical(theta, phi, radius):
    """load_blender.py:29-34: translate along z, rotate by phi about x and theta about y, swap to the Blender frame."""
    ph, th = phi / 180. * np.pi, theta / 180. * np.pi
    trans_t = torch.Tensor([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, radius], [0, 0, 0, 1]]).float()
    rot_phi = torch.Tensor([[1, 0, 0, 0], [0, np.cos(ph), -np.sin(ph), 0], [0, np.sin(ph), np.cos(ph), 0], [0, 0, 0, 1]]).float()
    rot_theta = torch.Tensor([[np.cos(th), 0, -np.sin(th), 0], [0, 1, 0, 0], [np.sin(th), 0, np.cos(th), 0], [0, 0, 0, 1]]).float()
    c2w = rot_theta @ (rot_phi @ trans_t)
    return torch.Tensor(np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]])) @ c2w


def _imread_rgba(fname):
    """What imageio.imread returns for these files: uint8 [H,W,4] RGBA (RGB for 3-channel files)."""
    import cv2
    img = cv2.imread(fname, cv2.IMREAD_UNCHANGED)
    if img is None:
        raise FileNotFoundError(fname)
    if img.ndim == 3 and img.shape[2] == 4:
        return img[..., [2, 1, 0, 3]]
    if img.ndim == 3:
        return img[..., ::-1]
    return img


def _half(imgs, H, W):
    import cv2
    out = np.zeros((imgs.shape[0], H, W, 4))                       # float64 like the reference (load_blender.py:94-102)
    for i, img in enumerate(imgs):
        out[i] = cv2.resize(img, (W, H), interpolation=cv2.INTER_AREA)
    return out


def load_blender_data(basedir, half_res=False, testskip=1, train_dir=None, workers=None):
    """load_blender.py:37-110.  Returns (imgs, poses, render_poses, [H, W, focal], i_split); with train_dir the first
    element is [train_imgs_from_train_dir, other_imgs] and i_split still counts the train frames first, exactly like the
    reference (run_nerf.py:574-596 consumes that pair)."""
    splits = ['train', 'val', 'test']
    metas = {}
    for s in splits:
        with open(os.path.join(basedir, 'transforms_{}.json'.format(s)), 'r') as fp:
            metas[s] = json.load(fp)
    jobs, poses_by_split = [], []
    for s in splits:
        skip = 1 if (s == 'train' or testskip == 0) else testskip
        names, poses = [], []
        for frame in metas[s]['frames'][::skip]:
            fname = os.path.join(basedir, frame['file_path'] + '.png')
            if s == 'train' and train_dir is not None:
                fname = os.path.join(train_dir, os.path.basename(fname))
            names.append(fname)
            poses.append(np.array(frame['transform_matrix']))
        jobs.append(names)
        poses_by_split.append(np.array(poses).astype(np.float32))
    # decode AND normalise on the pool: (uint8 / 255.) evaluated in float64 and rounded to float32 straight into the
    # split's output array is the reference's (np.array(imgs) / 255.).astype(np.float32) element for element, without
    # its float64 temporary and without a stacking copy (numpy and cv2 release the GIL inside all three steps)
    def _normalise(job):
        dst, img = job
        np.divide(img, 255., out=dst, dtype=np.float64, casting='unsafe')
    with ThreadPoolExecutor(max_workers=workers or min(32, os.cpu_count() or 1)) as pool:
        decoded = []
        for names in jobs:
            raw = list(pool.map(_imread_rgba, names))
            if not raw:
                decoded.append(np.zeros((0,), np.float32))
                continue
            arr = np.empty((len(raw),) + raw[0].shape, np.float32)
            list(pool.map(_normalise, [(arr[i], im) for i, im in enumerate(raw)]))
            decoded.append(arr)

    all_imgs, train_imgs, counts = [], [], [0]
    for s, imgs in zip(splits, decoded):                            # keep all 4 channels (RGBA)
        counts.append(counts[-1] + imgs.shape[0])
        if s == 'train' and train_dir is not None:
            train_imgs.append(imgs)
        else:
            all_imgs.append(imgs)
    i_split = [np.arange(counts[i], counts[i + 1]) for i in range(3)]
    imgs = np.concatenate(all_imgs, 0)
    poses = np.concatenate(poses_by_split, 0)
    train_np_imgs = np.concatenate(train_imgs, 0) if train_dir is not None else None

    H, W = imgs[0].shape[:2]
    camera_angle_x = float(metas['test']['camera_angle_x'])         # the reference reads the last split's meta
    focal = .5 * W / np.tan(.5 * camera_angle_x)
    render_poses = torch.stack([pose_spherical(angle, -30.0, 4.0) for angle in np.linspace(-180, 180, 40 + 1)[:-1]], 0)

    if half_res:
        H, W, focal = H // 2, W // 2, focal / 2.
        if train_dir is not None:
            train_np_imgs = _half(train_np_imgs, H, W)
        imgs = _half(imgs, H, W)
    if train_dir is not None:
        return [train_np_imgs, imgs], poses, render_poses, [H, W, focal], i_split
    return imgs, poses, render_poses, [H, W, focal], i_split


def white_background(images):
    """run_nerf.py:584-596 (`white_bkgd`): rgb * alpha + (1 - alpha), for an array or the [train, rest] pair."""
    if isinstance(images, (list, tuple)):
        return [white_background(im) for im in images]
    return images[..., :3] * images[..., -1:] + (1. - images[..., -1:])
