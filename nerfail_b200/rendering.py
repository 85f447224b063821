"""Host-side mirror of nerf-pytorch's render path with the reference's signatures.

Reference: Create_spatial_point_set/nerf_pytorch/run_nerf.py (render :69, batchify_rays :54, render_rays :308,
run_network :37, raw2outputs :262, create_nerf :178, render_path :137) and the pts_max variant in
Create_spatial_point_set/nerf_to_coord.py:70-135, :418-423.  The Python here only sequences kernel launches
of libnerfail_b200.so; it holds no arithmetic of its own on the hot path.
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch

from . import nerf, ops
from .nerf import NeRF, get_embedder
from .optim import Adam

DEBUG = False


def batchify(fn, chunk):
    """run_nerf.py:27-34."""
    if chunk is None:
        return fn

    def ret(inputs):
        return torch.cat([fn(inputs[i:i + chunk]) for i in range(0, inputs.shape[0], chunk)], 0)
    return ret


def _fusable(fn, embed_fn, embeddirs_fn) -> bool:
    return (isinstance(fn, NeRF) and fn.fused_supported()
            and isinstance(embed_fn, nerf.Embedder) and embed_fn.multires == 10
            and isinstance(embeddirs_fn, nerf.Embedder) and embeddirs_fn.multires == 4
            and nerf.mlp_precision() == "bf16" and not torch.is_grad_enabled())


def run_network(inputs, viewdirs, fn, embed_fn, embeddirs_fn, netchunk=1024 * 64):
    """run_nerf.py:37-51.  inputs [R,S,3], viewdirs [R,3] -> [R,S,4].

    Without autograd and with the reference architecture the whole call is ONE fused kernel (encoding, both
    concats and all 12 linear layers); otherwise the encodings are written once and the fp32 layer kernels run
    in netchunk pieces like the reference.
    """
    if viewdirs is not None and _fusable(fn, embed_fn, embeddirs_fn):
        return fn.fused().forward_points(inputs, viewdirs)
    S = inputs.shape[-2] if inputs.dim() >= 2 else 1
    flat = inputs.reshape(-1, inputs.shape[-1])
    wants_input_grad = torch.is_grad_enabled() and (flat.requires_grad or (viewdirs is not None and viewdirs.requires_grad))
    if isinstance(embed_fn, nerf.Embedder) and wants_input_grad:
        # pose / ray optimisation on the drop-in names: keep the gradient to the points and directions like the
        # reference's torch Embedder (ops.EmbedFn), at the price of a concat
        embedded = embed_fn.embed_grad(flat)
        if viewdirs is not None:
            embedded = torch.cat([embedded, embeddirs_fn.embed_grad(viewdirs.reshape(-1, 3), row_repeat=S)], -1)
    elif isinstance(embed_fn, nerf.Embedder):
        width = embed_fn.out_dim + (embeddirs_fn.out_dim if viewdirs is not None else 0)
        embedded = torch.empty((flat.shape[0], width), dtype=torch.float32, device=flat.device)
        embed_fn.embed(flat, embedded, 0)
        if viewdirs is not None:
            embeddirs_fn.embed(viewdirs.reshape(-1, 3), embedded, embed_fn.out_dim, row_repeat=S)
    else:   # i_embed == -1: identity embedding
        embedded = embed_fn(flat)
        if viewdirs is not None:
            dirs = viewdirs[:, None].expand(inputs.shape).reshape(-1, inputs.shape[-1])
            embedded = torch.cat([embedded, embeddirs_fn(dirs)], -1)
    outputs_flat = batchify(fn, netchunk)(embedded)
    return outputs_flat.reshape(list(inputs.shape[:-1]) + [outputs_flat.shape[-1]])


class NetworkQuery:
    """The `network_query_fn` built by create_nerf (run_nerf.py:201-204) as an object, so that render_rays can
    recognise it and hand rays + depths straight to the fused kernel (points formed in-kernel)."""

    def __init__(self, embed_fn, embeddirs_fn, netchunk):
        self.embed_fn, self.embeddirs_fn, self.netchunk = embed_fn, embeddirs_fn, netchunk

    def __call__(self, inputs, viewdirs, network_fn):
        return run_network(inputs, viewdirs, network_fn, embed_fn=self.embed_fn, embeddirs_fn=self.embeddirs_fn,
                           netchunk=self.netchunk)

    def from_rays(self, ray_batch, z_vals, network_fn):
        """raw [R,S,4] for pts = o + d*z (run_nerf.py:381) without materialising pts when the fused path applies."""
        if ray_batch.shape[-1] == 11 and _fusable(network_fn, self.embed_fn, self.embeddirs_fn):
            return network_fn.fused().forward_rays(ray_batch, z_vals)
        if (ray_batch.shape[-1] == 11 and torch.is_grad_enabled() and nerf.train_precision() == "bf16"
                and not ray_batch.requires_grad and not z_vals.requires_grad      # the fused backward reaches the parameters only
                and isinstance(network_fn, NeRF) and network_fn.fused_supported()
                and isinstance(self.embed_fn, nerf.Embedder) and self.embed_fn.multires == 10
                and isinstance(self.embeddirs_fn, nerf.Embedder) and self.embeddirs_fn.multires == 4):
            return network_fn.forward_rays_train(ray_batch, z_vals)
        rays_o, rays_d = ray_batch[:, 0:3], ray_batch[:, 3:6]
        viewdirs = ray_batch[:, -3:] if ray_batch.shape[-1] > 8 else None
        pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]
        return self(pts, viewdirs, network_fn)


def raw2outputs(raw, z_vals, rays_d, raw_noise_std=0, white_bkgd=False, pytest=False):
    """run_nerf.py:262-305 -> (rgb_map, disp_map, acc_map, weights, depth_map); differentiable w.r.t. raw."""
    noise = None
    if raw_noise_std > 0.0:
        noise = torch.randn(raw[..., 3].shape, device=raw.device) * raw_noise_std
        if pytest:   # the reference's deterministic hook (:288-291) draws *uniform* numbers here
            np.random.seed(0)
            noise = torch.tensor(np.random.rand(*list(raw[..., 3].shape)) * raw_noise_std, dtype=torch.float32,
                                 device=raw.device)
    if raw.requires_grad and torch.is_grad_enabled():
        return ops.CompositeFn.apply(raw, z_vals, rays_d, noise, white_bkgd)
    return ops.composite_fwd(raw, z_vals, rays_d, noise, white_bkgd)


def render_rays(ray_batch, network_fn, network_query_fn, N_samples, retraw=False, lindisp=False, perturb=0.,
                N_importance=0, network_fine=None, white_bkgd=False, raw_noise_std=0., verbose=False, pytest=False,
                with_pts_max=False):
    """run_nerf.py:308-418 (and nerf_to_coord.py:320-433 when with_pts_max=True)."""
    ray_batch = ray_batch.float().contiguous()
    N_rays = ray_batch.shape[0]
    # Whole-batch fast path: without autograd, noise and the raw output nothing but kernel launches is left of this
    # function, so they are sequenced by ONE C call (nfb_render_rays_fwd); NERFAIL_B200_RENDER_RAYS=ops keeps the per-op path.
    if (not torch.is_grad_enabled() and not retraw and raw_noise_std == 0. and not pytest and ray_batch.shape[-1] == 11
            and isinstance(network_query_fn, NetworkQuery)
            and _fusable(network_fn, network_query_fn.embed_fn, network_query_fn.embeddirs_fn)
            and (network_fine is None or _fusable(network_fine, network_query_fn.embed_fn, network_query_fn.embeddirs_fn))
            and 3 <= N_samples <= 128 and N_samples + N_importance <= 512
            and os.environ.get("NERFAIL_B200_RENDER_RAYS", "fused") != "ops"):
        dev = ray_batch.device
        t_r = u_r = rng = None
        if perturb != 0.:
            # stratified draws: Philox inside the kernels (no [R,64] + [R,128] HBM tensors, no torch.rand launches) unless
            # NERFAIL_B200_RNG=torch asks for torch's generator (what the per-op path and autograd use).  The reference
            # jitters the coarse depths for perturb > 0 (:365) and draws u for perturb != 0 (:393, det = perturb == 0).
            if os.environ.get("NERFAIL_B200_RNG", "philox") == "torch" or perturb < 0.:
                t_r = torch.rand((N_rays, N_samples), device=dev) if perturb > 0. else None
                u_r = torch.rand((N_rays, N_importance), device=dev) if N_importance > 0 else None
            else:
                rng = ops.next_philox()
        fine = network_fine.fused() if network_fine is not None else None
        return ops.render_rays_fused(network_fn.fused(), fine, ray_batch, N_samples, N_importance, lindisp, white_bkgd,
                                     t_r, u_r, want_pts_max=with_pts_max, rng=rng)
    t_rand = None
    if perturb > 0.:
        if pytest:
            np.random.seed(0)
            t_rand = torch.tensor(np.random.rand(N_rays, N_samples), dtype=torch.float32, device=ray_batch.device)
        else:
            t_rand = torch.rand((N_rays, N_samples), device=ray_batch.device)
    # Whole-batch training path: under autograd with the tensor-core training kernels, no noise and both networks present,
    # the forward and the backward of this function are ONE C call each (nfb_render_rays_train_fwd / nfb_render_rays_bwd):
    # the same kernels in the same order as the per-op path below, which NERFAIL_B200_RENDER_RAYS=ops keeps.
    if (torch.is_grad_enabled() and raw_noise_std == 0. and N_importance > 0 and network_fine is not None and not with_pts_max
            and ray_batch.shape[-1] == 11 and not ray_batch.requires_grad and nerf.train_precision() == "bf16"
            and isinstance(network_query_fn, NetworkQuery)
            and isinstance(network_query_fn.embed_fn, nerf.Embedder) and network_query_fn.embed_fn.multires == 10
            and isinstance(network_query_fn.embeddirs_fn, nerf.Embedder) and network_query_fn.embeddirs_fn.multires == 4
            and isinstance(network_fn, NeRF) and network_fn.fused_supported()
            and isinstance(network_fine, NeRF) and network_fine.fused_supported() and network_fine is not network_fn
            and 3 <= N_samples <= 128 and N_samples + N_importance <= 512 and N_rays > 0
            and os.environ.get("NERFAIL_B200_RENDER_RAYS", "fused") != "ops"):
        u = None
        if perturb != 0.:
            if pytest:
                np.random.seed(0)
                u = torch.tensor(np.random.rand(N_rays, N_importance), dtype=torch.float32, device=ray_batch.device)
            else:
                u = torch.rand((N_rays, N_importance), device=ray_batch.device)
        rgb_map, disp_map, acc_map, rgb0, disp0, acc0, z_std, raw = ops.RenderRaysTrainFn.apply(
            network_fn, network_fine, ray_batch, N_samples, N_importance, lindisp, white_bkgd, t_rand, u,
            *network_fn.ordered_params(), *network_fine.ordered_params())
        ret = {'rgb_map': rgb_map, 'disp_map': disp_map, 'acc_map': acc_map}
        if retraw:
            ret['raw'] = raw
        ret.update(rgb0=rgb0, disp0=disp0, acc0=acc0, z_std=z_std)
        return ret
    z_vals = ops.coarse_z(ray_batch, N_samples, lindisp, t_rand)

    query_rays = network_query_fn.from_rays if isinstance(network_query_fn, NetworkQuery) else None

    def query(z, net):
        if query_rays is not None:
            return query_rays(ray_batch, z, net)
        rays_o, rays_d = ray_batch[:, 0:3], ray_batch[:, 3:6]
        viewdirs = ray_batch[:, -3:] if ray_batch.shape[-1] > 8 else None
        pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]
        return network_query_fn(pts, viewdirs, net)

    def composite(raw, z, want_pts_max):
        noise = None
        if raw_noise_std > 0.:
            noise = torch.randn(raw[..., 3].shape, device=raw.device) * raw_noise_std
            if pytest:
                np.random.seed(0)
                noise = torch.tensor(np.random.rand(*list(raw[..., 3].shape)) * raw_noise_std, dtype=torch.float32,
                                     device=raw.device)
        if raw.requires_grad and torch.is_grad_enabled():
            out = ops.CompositeFn.apply(raw, z, ray_batch, noise, white_bkgd)
            pm = None
            if want_pts_max:
                with torch.no_grad():
                    pm = ops.composite_fwd(raw.detach(), z, ray_batch, noise, white_bkgd, want_pts_max=True)[5]
            return out + (pm,)
        out = ops.composite_fwd(raw, z, ray_batch, noise, white_bkgd, want_pts_max=want_pts_max)
        return out if want_pts_max else out + (None,)

    raw = query(z_vals, network_fn)
    rgb_map, disp_map, acc_map, weights, depth_map, pts_max = composite(raw, z_vals, with_pts_max and N_importance <= 0)

    if N_importance > 0:
        rgb_map_0, disp_map_0, acc_map_0 = rgb_map, disp_map, acc_map
        u = None
        if perturb != 0.:
            if pytest:
                np.random.seed(0)
                u = torch.tensor(np.random.rand(N_rays, N_importance), dtype=torch.float32, device=ray_batch.device)
            else:
                u = torch.rand((N_rays, N_importance), device=ray_batch.device)
        # z_mid, sample_pdf(weights[...,1:-1]), detach, sort(cat) and z_std in one kernel (:392-396, :412)
        if 3 <= N_samples <= 128 and N_samples + N_importance <= 512:
            z_vals, z_samples, z_std = ops.hierarchical(z_vals, weights.detach(), N_importance, u)
        else:       # outside the fused kernel's shapes: the reference's own sequence (:392-396, :412) on the sample_pdf kernel
            z_mid = .5 * (z_vals[..., 1:] + z_vals[..., :-1])
            z_samples = ops.sample_pdf(z_mid, weights.detach()[..., 1:-1].contiguous(), N_importance, u)
            z_std = torch.std(z_samples, dim=-1, unbiased=False)
            z_vals, _ = torch.sort(torch.cat([z_vals, z_samples], -1), -1)
        run_fn = network_fn if network_fine is None else network_fine
        raw = query(z_vals, run_fn)
        rgb_map, disp_map, acc_map, weights, depth_map, pts_max = composite(raw, z_vals, with_pts_max)

    ret = {'rgb_map': rgb_map, 'disp_map': disp_map, 'acc_map': acc_map}
    if with_pts_max:
        ret['pts_max'] = pts_max
    if retraw:
        ret['raw'] = raw
    if N_importance > 0:
        ret['rgb0'] = rgb_map_0
        ret['disp0'] = disp_map_0
        ret['acc0'] = acc_map_0
        ret['z_std'] = z_std

    if DEBUG:
        for k in ret:
            if torch.isnan(ret[k]).any() or torch.isinf(ret[k]).any():
                print(f"! [Numerical Error] {k} contains nan or inf.")
    return ret


# Rays are independent, so the `chunk` argument bounds memory only ("Does not affect final results",
# run_nerf.py:78-79).  A B200 has room for far more than the reference's 1024*32 rays per pass and the fused
# MLP wants thousands of 128-sample tiles per launch, so consecutive chunks are coalesced up to this many
# rays per kernel pass unless NERFAIL_B200_STRICT_CHUNK=1 asks for the reference's literal chunking.
MAX_RAYS_PER_PASS = int(os.environ.get("NERFAIL_B200_RAYS_PER_PASS", 1 << 18))


def batchify_rays(rays_flat, chunk=1024 * 32, **kwargs):
    """run_nerf.py:54-66."""
    if os.environ.get("NERFAIL_B200_STRICT_CHUNK", "0") != "1" and not torch.is_grad_enabled():
        chunk = max(chunk, MAX_RAYS_PER_PASS)
    all_ret = {}
    for i in range(0, rays_flat.shape[0], chunk):
        ret = render_rays(rays_flat[i:i + chunk], **kwargs)
        for k in ret:
            all_ret.setdefault(k, []).append(ret[k])
    return {k: (v[0] if len(v) == 1 else torch.cat(v, 0)) for k, v in all_ret.items()}


def render(H, W, K, chunk=1024 * 32, rays=None, c2w=None, ndc=True, near=0., far=1., use_viewdirs=False,
           c2w_staticcam=None, with_pts_max=False, **kwargs):
    """run_nerf.py:69-134.  Returns [rgb_map, disp_map, acc_map, extras]; with_pts_max=True inserts pts_max
    before extras like nerf_to_coord.py:132-135."""
    if c2w is not None and not ndc and c2w_staticcam is None and use_viewdirs:
        # full-image fast path: one kernel emits the [H*W,11] batch (get_rays + viewdirs + near/far)
        dev = c2w.device if isinstance(c2w, torch.Tensor) and c2w.is_cuda else torch.device("cuda")
        rays_b = ops.get_ray_batch(H, W, K, c2w, near, far, device=dev)
        sh = (H, W, 3)
    elif (c2w is None and not ndc and c2w_staticcam is None and use_viewdirs and isinstance(rays, torch.Tensor) and rays.is_cuda
          and rays.dim() == 3 and rays.shape[0] == 2 and rays.shape[-1] == 3 and not rays.requires_grad):
        # the training step's call (run_nerf.py:776: rays=batch_rays [2,N,3]): one kernel builds the [N,11] batch
        rays_b = ops.rays_from_batch(rays[0], rays[1], near, far)
        sh = (rays.shape[1], 3)
    else:
        if c2w is not None:
            rays_o, rays_d = nerf.get_rays(H, W, K, c2w)
        else:
            rays_o, rays_d = rays
        viewdirs = None
        if use_viewdirs:
            viewdirs = rays_d
            if c2w_staticcam is not None:
                rays_o, rays_d = nerf.get_rays(H, W, K, c2w_staticcam)
            viewdirs = viewdirs / torch.norm(viewdirs, dim=-1, keepdim=True)
            viewdirs = torch.reshape(viewdirs, [-1, 3]).float()
        sh = rays_d.shape
        if ndc:
            rays_o, rays_d = nerf.ndc_rays(H, W, K[0][0], 1., rays_o, rays_d)
        rays_o = torch.reshape(rays_o, [-1, 3]).float()
        rays_d = torch.reshape(rays_d, [-1, 3]).float()
        near_t, far_t = near * torch.ones_like(rays_d[..., :1]), far * torch.ones_like(rays_d[..., :1])
        rays_b = torch.cat([rays_o, rays_d, near_t, far_t], -1)
        if use_viewdirs:
            rays_b = torch.cat([rays_b, viewdirs], -1)
        if not rays_b.is_cuda:
            raise RuntimeError("nerfail_b200.render needs CUDA rays")

    all_ret = batchify_rays(rays_b, chunk, with_pts_max=with_pts_max, **kwargs)
    for k in all_ret:
        k_sh = list(sh[:-1]) + list(all_ret[k].shape[1:])
        all_ret[k] = torch.reshape(all_ret[k], k_sh)

    k_extract = ['rgb_map', 'disp_map', 'acc_map'] + (['pts_max'] if with_pts_max else [])
    ret_list = [all_ret[k] for k in k_extract]
    ret_dict = {k: all_ret[k] for k in all_ret if k not in k_extract}
    return ret_list + [ret_dict]


class _ViewSink:
    """Takes finished views off the GPU without stalling the launch loop: device->host copies into a ring of pinned
    buffers on a side stream, PNG / .npy encoding on a writer thread (cv2 and numpy release the GIL).  The reference
    does `.cpu().numpy()` + imageio.imwrite inline per view (run_nerf.py:155-169, nerf_to_coord.py:155-173)."""

    def __init__(self, n_views, H, W, with_pts, savedir, device, depth=3, nets=()):
        import queue
        import threading
        self.savedir, self.with_pts = savedir, with_pts
        self.nets = [n for n in nets if isinstance(n, NeRF)]      # polled for a barrier time-out once per finished view
        self.rgbs = np.empty((n_views, H, W, 3), np.float32)
        self.disps = np.empty((n_views, H, W), np.float32)
        self.pts = np.empty((n_views, H, W, 3), np.float32) if with_pts else None
        self.stream = torch.cuda.Stream(device=device)
        self.slots = [{"rgb": torch.empty((H, W, 3), dtype=torch.float32).pin_memory(),
                       "disp": torch.empty((H, W), dtype=torch.float32).pin_memory(),
                       "pts": torch.empty((H, W, 3), dtype=torch.float32).pin_memory() if with_pts else None,
                       "free": threading.Event()} for _ in range(depth)]
        for sl in self.slots:
            sl["free"].set()
        self.q = queue.Queue()
        self.err = None
        self.thread = threading.Thread(target=self._drain, daemon=True)
        self.thread.start()

    def put(self, k, name, rgb, disp, pts):
        sl = self.slots[k % len(self.slots)]
        sl["free"].wait()
        sl["free"].clear()
        self.stream.wait_stream(torch.cuda.current_stream(rgb.device))
        with torch.cuda.stream(self.stream):
            sl["rgb"].copy_(rgb, non_blocking=True)
            sl["disp"].copy_(disp, non_blocking=True)
            if self.with_pts:
                sl["pts"].copy_(pts, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.stream)
        self.q.put((k, name, sl, done, (rgb, disp, pts)))         # the device tensors stay referenced until copied

    def _drain(self):
        while True:
            job = self.q.get()
            if job is None:
                return
            k, name, sl, done, _keep = job
            try:
                done.synchronize()
                for n in self.nets:                                    # the view's kernels have finished: one host read each
                    if n._fused is not None:
                        n._fused.poll()
                self.rgbs[k] = sl["rgb"].numpy()
                self.disps[k] = sl["disp"].numpy()
                if self.with_pts:
                    self.pts[k] = sl["pts"].numpy()
                sl["free"].set()
                if self.savedir is not None:
                    import cv2
                    cv2.imwrite(os.path.join(self.savedir, name + '.png'), nerf.to8b(self.rgbs[k])[..., ::-1])
                    if self.with_pts:
                        np.save(os.path.join(self.savedir, name + '.npy'), self.pts[k])
            except Exception as e:                                     # surfaced by close()
                self.err = e
                sl["free"].set()

    def close(self):
        self.q.put(None)
        self.thread.join()
        if self.err is not None:
            raise self.err


def render_path(render_poses, hwf, K, chunk, render_kwargs, gt_imgs=None, savedir=None, render_factor=0,
                with_pts_max=False, view_ids=None):
    """run_nerf.py:137-175 / nerf_to_coord.py:138-180: render every pose, optionally saving PNG (+ pts_max .npy).
    The launch loop never waits for the host: finished views leave through _ViewSink.  view_ids names the files when
    the poses are one rank's shard of a sweep (render_sweep); default 0, 1, ... like the reference."""
    H, W, focal = hwf
    if render_factor != 0:
        H, W, focal = H // render_factor, W // render_factor, focal / render_factor
    n = len(render_poses)
    ids = list(range(n)) if view_ids is None else list(view_ids)
    dev = None
    sink = None
    try:
        for k, c2w in enumerate(render_poses):
            t0 = time.time()
            out = render(H, W, K, chunk=chunk, c2w=c2w[:3, :4], with_pts_max=with_pts_max, **render_kwargs)
            if sink is None:
                dev = out[0].device
                sink = _ViewSink(n, H, W, with_pts_max, savedir, dev,
                                 nets=(render_kwargs.get("network_fn"), render_kwargs.get("network_fine")))
            sink.put(k, '{:03d}'.format(ids[k]), out[0], out[1], out[3] if with_pts_max else None)
            if DEBUG:
                print(k, time.time() - t0)
    finally:
        if sink is not None:
            sink.close()
    if sink is None:
        z = np.zeros((0, H, W, 3), np.float32)
        return (z, z[..., 0], z) if with_pts_max else (z, z[..., 0])
    return (sink.rgbs, sink.disps, sink.pts) if with_pts_max else (sink.rgbs, sink.disps)


def render_sweep(render_poses, hwf, K, chunk, render_kwargs, savedir=None, with_pts_max=False, rank=None,
                 world_size=None):
    """The novel-view sweep of nerf_render_only.py / nerf_to_coord.py (:619-648) sharded by view: rank r renders the
    poses i with i % world_size == r (views are independent, no collective) and names its files by the global view
    index, so the ranks of a node fill one directory exactly like the reference's single process.  Returns
    (view_ids, rgbs, disps[, pts_max]) of this rank."""
    from . import dist
    r, w = dist.world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    ids = dist.shard_views(len(render_poses), rank, world_size)
    out = render_path([render_poses[i] for i in ids], hwf, K, chunk, render_kwargs, savedir=savedir,
                      with_pts_max=with_pts_max, view_ids=ids)
    return (ids,) + tuple(out)


def create_nerf(args, device=None):
    """run_nerf.py:178-259: embedders, coarse/fine NeRF, Adam, checkpoint reload, render kwargs."""
    device = torch.device(device if device is not None else "cuda")
    embed_fn, input_ch = get_embedder(args.multires, args.i_embed)
    input_ch_views, embeddirs_fn = 0, None
    if args.use_viewdirs:
        embeddirs_fn, input_ch_views = get_embedder(args.multires_views, args.i_embed)
    output_ch = 5 if args.N_importance > 0 else 4
    skips = [4]
    model = NeRF(D=args.netdepth, W=args.netwidth, input_ch=input_ch, output_ch=output_ch, skips=skips,
                 input_ch_views=input_ch_views, use_viewdirs=args.use_viewdirs).to(device)
    grad_vars = list(model.parameters())
    model_fine = None
    if args.N_importance > 0:
        model_fine = NeRF(D=args.netdepth_fine, W=args.netwidth_fine, input_ch=input_ch, output_ch=output_ch,
                          skips=skips, input_ch_views=input_ch_views, use_viewdirs=args.use_viewdirs).to(device)
        grad_vars += list(model_fine.parameters())

    network_query_fn = NetworkQuery(embed_fn, embeddirs_fn, args.netchunk)
    optimizer = Adam(params=grad_vars, lr=args.lrate, betas=(0.9, 0.999))       # fused multi-tensor kernel, torch.optim.Adam state_dict

    start = 0
    basedir, expname = getattr(args, 'basedir', None), getattr(args, 'expname', None)
    ckpts = []
    if getattr(args, 'ft_path', None) is not None and args.ft_path != 'None':
        ckpts = [args.ft_path]
    elif basedir is not None and expname is not None and os.path.isdir(os.path.join(basedir, expname)):
        ckpts = [os.path.join(basedir, expname, f) for f in sorted(os.listdir(os.path.join(basedir, expname)))
                 if 'tar' in f]
    if len(ckpts) > 0 and not getattr(args, 'no_reload', False):
        ckpt = torch.load(ckpts[-1], map_location=device)
        start = ckpt['global_step']
        optimizer.load_state_dict(ckpt['optimizer_state_dict'])
        model.load_state_dict(ckpt['network_fn_state_dict'])
        if model_fine is not None:
            model_fine.load_state_dict(ckpt['network_fine_state_dict'])

    render_kwargs_train = {
        'network_query_fn': network_query_fn,
        'perturb': args.perturb,
        'N_importance': args.N_importance,
        'network_fine': model_fine,
        'N_samples': args.N_samples,
        'network_fn': model,
        'use_viewdirs': args.use_viewdirs,
        'white_bkgd': args.white_bkgd,
        'raw_noise_std': args.raw_noise_std,
    }
    if args.dataset_type != 'llff' or args.no_ndc:
        render_kwargs_train['ndc'] = False
        render_kwargs_train['lindisp'] = args.lindisp
    render_kwargs_test = {k: render_kwargs_train[k] for k in render_kwargs_train}
    render_kwargs_test['perturb'] = False
    render_kwargs_test['raw_noise_std'] = 0.
    return render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer
