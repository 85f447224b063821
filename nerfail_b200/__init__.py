"""nerfail_b200 — B200-native (sm_100a) kernels behind NeRFail's differentiable-rendering API.

Drop-in names (same signatures as the reference):
    render, render_rays, run_network, raw2outputs, sample_pdf, batchify, batchify_rays, create_nerf,
    NeRF, get_embedder, get_rays, gauss_net, create_gauss_w
Importing the package does not need a GPU; calling any op does (there is no CPU fallback).
"""
from .nerf import NeRF, Embedder, get_embedder, get_rays, get_rays_np, ndc_rays, sample_pdf, img2mse, mse2psnr, to8b
from .rendering import (render, render_rays, run_network, raw2outputs, batchify, batchify_rays, render_path, render_sweep,
                     create_nerf, NetworkQuery)
from .gauss import gauss_net, create_gauss_w, knn_index_and_dist
from .optim import Adam, decayed_lrate, set_lrate
from .pipeline import SpatialPointSet, render_points, build_attack_inputs, knn_sweep, AttackImageSink
from .data import load_blender_data, pose_spherical, white_background
from .train import sample_ray_batch, train_step, make_train_stepper, save_checkpoint, checkpoint_path, precrop_window

__all__ = [
    "NeRF", "Embedder", "get_embedder", "get_rays", "get_rays_np", "ndc_rays", "sample_pdf", "img2mse", "mse2psnr",
    "to8b", "render", "render_rays", "run_network", "raw2outputs", "batchify", "batchify_rays", "render_path",
    "render_sweep", "create_nerf", "NetworkQuery", "gauss_net", "create_gauss_w", "knn_index_and_dist", "Adam", "decayed_lrate", "set_lrate", "SpatialPointSet", "render_points", "build_attack_inputs", "knn_sweep", "AttackImageSink",
    "load_blender_data", "pose_spherical", "white_background", "sample_ray_batch", "train_step", "make_train_stepper", "save_checkpoint", "checkpoint_path", "precrop_window",
]
__version__ = "0.1.0"
