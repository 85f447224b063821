"""On-device index / weight pipeline.

The reference builds the inputs of `gauss_net` through three scripts and the disk:
    nerf_to_coord.py:166-173        render every view, save pts_max as NNN.npy  (float32 [H,W,3])
    create_index_and_dist.py:96-163 load the P base views + every view, 8-NN, save index_and_dist/<i>.pth  (float32 [2,H,W,8])
    tools/dist_to_weight.py:80-97   create_gauss_w (c = 0.02), save index_and_weight/<i>.pth              (float32 [2,H,W,8])
and `MyDataset.py:201` torch.load()s one 41 MB file per view and iteration.  Here the same tensors are produced on the
GPU and handed to `gauss_net` directly: render(with_pts_max=True) -> KnnGrid.query -> gauss_weights.  The readers and
writers below keep the reference's file formats for interchange (files written here load in the reference and vice versa).
"""
from __future__ import annotations

import os
from typing import Iterable, Optional, Sequence

import numpy as np
import torch

from . import ops
from .rendering import render

GAUSS_C = 0.02          # tools/dist_to_weight.py:80


# ---- file formats of the reference ------------------------------------------------------------------------------
def save_points_npy(path: str, pts_hw3: torch.Tensor) -> None:
    """nerf_to_coord.py:172-173: np.save of the float32 [H,W,3] arg-max-weight points of one view."""
    np.save(path, pts_hw3.detach().to(torch.float32).cpu().numpy())


def load_points_npy(path: str, device=None) -> torch.Tensor:
    """create_index_and_dist.py:62-66 (np.load + torch.from_numpy + .to(device))."""
    return torch.from_numpy(np.load(path).astype(np.float32, copy=False)).to(device if device is not None else "cuda")


def save_index_and_dist(path: str, dist_idx: torch.Tensor) -> None:
    """create_index_and_dist.py:148-163: torch.save of float32 [2,H,W,8] = cat([dist, idx]); also the format of
    tools/dist_to_weight.py:95-97 for cat([weight, idx])."""
    if dist_idx.dim() != 4 or dist_idx.shape[0] != 2 or dist_idx.shape[-1] != 8:
        raise ValueError("expected a [2,H,W,8] tensor")
    torch.save(dist_idx.detach().to(torch.float32), path)


def load_index_and_dist(path: str, device=None) -> torch.Tensor:
    """MyDataset.py:201: torch.load of a [2,H,W,8] float32 tensor."""
    return torch.load(path, map_location=device if device is not None else "cuda")


save_index_and_weight, load_index_and_weight = save_index_and_dist, load_index_and_dist


# ---- the pipeline -------------------------------------------------------------------------------------------------
class SpatialPointSet:
    """The P base views' points (`test_index_change_tensor`, create_index_and_dist.py:96-108: the [P,H,W,3] pts_max of
    the views that carry the perturbation, flattened view-major) with their 8-NN grid.

    index_and_dist(view_pts)   == one iteration of create_index_and_dist.py:110-151       -> float32 [2,H,W,8]
    index_and_weight(view_pts) == + create_gauss_w (tools/dist_to_weight.py)               -> float32 [2,H,W,8]
    Row index r of `spatial_rgb.view(-1, 4)` (GaussNet.py:53) is the candidate index stored in channel 1.
    """

    def __init__(self, base_points: torch.Tensor, c: float = GAUSS_C):
        if base_points.dim() != 4 or base_points.shape[-1] != 3:
            raise ValueError("base_points must be [P,H,W,3]")
        if not base_points.is_cuda:
            raise RuntimeError("nerfail_b200.SpatialPointSet runs on CUDA only")
        self.shape = tuple(base_points.shape[:3])
        self.c = float(c)
        self.points = base_points.detach().to(torch.float32).reshape(-1, 3).contiguous()
        self.grid = ops.KnnGrid(self.points)

    @classmethod
    def from_npy(cls, paths: Sequence[str], device=None, c: float = GAUSS_C) -> "SpatialPointSet":
        return cls(torch.stack([load_points_npy(p, device) for p in paths], 0), c)

    def index_and_dist(self, view_points_hw3: torch.Tensor) -> torch.Tensor:
        return self.grid.query_dist_idx(view_points_hw3)

    def index_and_weight(self, view_points_hw3: torch.Tensor) -> torch.Tensor:
        di = self.index_and_dist(view_points_hw3)
        return ops.gauss_weights(di.unsqueeze(0), self.c)[0]

    def index_and_weight_batch(self, views_points: Iterable[torch.Tensor]) -> torch.Tensor:
        """[B,2,H,W,8] for a batch of views — what the attack loop feeds gauss_net.forward as weight_and_index_list."""
        di = torch.stack([self.index_and_dist(v) for v in views_points], 0)
        return ops.gauss_weights(di, self.c)


def render_points(H: int, W: int, K, c2w, chunk: int = 1024 * 32, **render_kwargs) -> torch.Tensor:
    """pts_max [H,W,3] of one view (nerf_to_coord.py:132-135, :418-423) without leaving the device."""
    with torch.no_grad():
        out = render(H, W, K, chunk=chunk, c2w=c2w, with_pts_max=True, **render_kwargs)
    return out[3]


def build_attack_inputs(H: int, W: int, K, base_poses, view_poses, render_kwargs: dict, chunk: int = 1024 * 32,
                        c: float = GAUSS_C, save_dir: Optional[str] = None):
    """Everything the three reference scripts produce for one scene, on the device:
    returns (SpatialPointSet of the P base views, weight_and_index_list [V,2,H,W,8] of `view_poses`).
    With save_dir the intermediate files are also written in the reference's layout (coords/NNN.npy,
    index_and_dist/<i>.pth, index_and_weight/<i>.pth)."""
    base = torch.stack([render_points(H, W, K, p[:3, :4], chunk, **render_kwargs) for p in base_poses], 0)
    sps = SpatialPointSet(base, c)
    out = []
    if save_dir is not None:
        for sub in ("coords", "index_and_dist", "index_and_weight"):
            os.makedirs(os.path.join(save_dir, sub), exist_ok=True)
    for i, pose in enumerate(view_poses):
        pts = render_points(H, W, K, pose[:3, :4], chunk, **render_kwargs)
        di = sps.index_and_dist(pts)
        iw = ops.gauss_weights(di.unsqueeze(0), sps.c)[0]
        out.append(iw)
        if save_dir is not None:
            save_points_npy(os.path.join(save_dir, "coords", "{:03d}.npy".format(i)), pts)
            save_index_and_dist(os.path.join(save_dir, "index_and_dist", f"{i}.pth"), di)
            save_index_and_weight(os.path.join(save_dir, "index_and_weight", f"{i}.pth"), iw)
    return sps, torch.stack(out, 0)


# ---- asynchronous file output shared by the sweep drivers ------------------------------------------------------
class _AsyncWriter:
    """Device results leave through a ring of pinned host buffers on a side stream; encoding / torch.save / cv2.imwrite run
    on a writer thread (they release the GIL), so the launch loop never waits for the disk.  `put(tensor, fn)` copies the
    tensor and later calls fn(host_numpy_or_tensor) on the thread; close() joins and re-raises the first error."""

    def __init__(self, device, depth: int = 4):
        import queue
        import threading
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.depth = depth
        self.free = threading.Semaphore(depth)
        self.q = queue.Queue()
        self.err = None
        self.thread = threading.Thread(target=self._drain, daemon=True)
        self.thread.start()

    def put(self, t: torch.Tensor, fn) -> None:
        self.free.acquire()
        host = torch.empty(t.shape, dtype=t.dtype).pin_memory()
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            host.copy_(t, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.stream)
        self.q.put((host, done, fn, t))                  # t stays referenced until the copy has finished

    def _drain(self):
        while True:
            job = self.q.get()
            if job is None:
                return
            host, done, fn, _keep = job
            try:
                done.synchronize()
                fn(host)
            except Exception as e:                       # surfaced by close()
                if self.err is None:
                    self.err = e
            finally:
                self.free.release()

    def close(self):
        self.q.put(None)
        self.thread.join()
        if self.err is not None:
            raise self.err


def knn_sweep(view_points: Sequence, base, out_dir: Optional[str] = None, rank: Optional[int] = None,
              world_size: Optional[int] = None, c: float = GAUSS_C, keep: bool = True):
    """The 8-NN precompute over all views of a data set (create_index_and_dist.py:110-163, then tools/dist_to_weight.py:80-97),
    SHARDED BY QUERY VIEW: rank r answers the views i with i % world_size == r against the replicated base points (23 MB for
    P = 3) — views are independent, there is no collective.  `view_points[i]` is a [H,W,3] tensor or the path
    of the view's NNN.npy; `base` a SpatialPointSet, a [P,H,W,3] tensor, or the list of the P base views' .npy paths.
    With out_dir every rank writes index_and_dist/<i>.pth and index_and_weight/<i>.pth in the reference's format through an
    asynchronous writer, the ranks of a node filling one directory like the reference's single process.
    Returns [(i, index_and_weight [2,H,W,8])] of this rank (empty tensors list if keep=False)."""
    from . import dist as nd
    r, w = nd.world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    dev = torch.device("cuda", torch.cuda.current_device())
    if isinstance(base, SpatialPointSet):
        sps = base
    elif isinstance(base, torch.Tensor):
        sps = SpatialPointSet(base.to(dev), c)
    else:
        sps = SpatialPointSet.from_npy(list(base), dev, c)
    writer = None
    if out_dir is not None:
        for sub in ("index_and_dist", "index_and_weight"):
            os.makedirs(os.path.join(out_dir, sub), exist_ok=True)
        writer = _AsyncWriter(dev)
    out = []
    try:
        for i in nd.shard_views(len(view_points), rank, world_size):
            v = view_points[i]
            pts = load_points_npy(v, dev) if isinstance(v, (str, os.PathLike)) else v.to(dev, torch.float32)
            di = sps.index_and_dist(pts)
            iw = ops.gauss_weights(di.unsqueeze(0), sps.c)[0]
            if writer is not None:
                writer.put(di, lambda h, p=os.path.join(out_dir, "index_and_dist", f"{i}.pth"): torch.save(h.clone(), p))
                writer.put(iw, lambda h, p=os.path.join(out_dir, "index_and_weight", f"{i}.pth"): torch.save(h.clone(), p))
            if keep:
                out.append((i, iw))
    finally:
        if writer is not None:
            writer.close()
    return out


class AttackImageSink:
    """The image output of the attack's last epoch (attack_NeRFail_S.py:394-403: per view `cv2.imwrite` of the perturbed RGBA
    image x_rgba, of the un-composited perturbation x and of the original, each a blocking `.cpu().detach().numpy()` + PNG
    encode on the attack's only thread) as an asynchronous sink: the float images are converted to the uint8 PNG payload on
    the device with cv2's own rule for float input (saturate_cast: round half to even, clamp to [0, 255]), cross PCIe at a
    quarter of the size, and are encoded on the writer thread.  Files are byte-identical to the reference's."""

    def __init__(self, device=None, depth: int = 6):
        dev = torch.device(device if device is not None else ("cuda", torch.cuda.current_device()))
        self.writer = _AsyncWriter(dev, depth)

    @staticmethod
    def _payload(img: torch.Tensor) -> torch.Tensor:
        if img.dtype == torch.uint8:
            return img.contiguous()
        return torch.round(img.detach().float()).clamp_(0, 255).to(torch.uint8)

    def put(self, x_rgba: torch.Tensor, x: torch.Tensor, ori_img: torch.Tensor, img_names: Sequence[str],
            mask_names: Sequence[str]) -> None:
        """One batch: x_rgba / x / ori_img [B,H,W,4]; img_names / mask_names as the reference's dataset yields them."""
        import cv2
        for b, (name, mname) in enumerate(zip(img_names, mask_names)):
            for t, path in ((x_rgba[b], name), (x[b], mname), (ori_img[b], name.replace(".png", "_ori.png"))):
                self.writer.put(self._payload(t), lambda h, p=path: cv2.imwrite(p, h.numpy()))

    def close(self) -> None:
        self.writer.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
