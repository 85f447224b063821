#!/usr/bin/env python
"""bench.py — render rays/s of the NeRFail differentiable-rendering hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): one 800x800 synthetic Blender-style view, lego config (64 coarse + 128
fine samples, 8x256 MLP, multires 10/4, white background), rendered through render()'s kernels.  One "step" =
one full view = 640 000 rays.  With N > 1 every rank renders its own view per step (views are independent —
run_nerf.py:151-154; weak scaling, no data-path collective) and `value` is the sum over ranks divided by the
slowest rank's device time.

Keys beyond the base contract:
  roofline      the fused MLP kernel (tensor-bound): algorithmic FLOP/launch / CUDA-event duration vs the measured
                bf16 peak of MEASURED_PEAKS.json (sustained figure — the kernel is timed inside a long step)
                roofline.traffic = dram bytes per launch from the committed ncu capture (profiles/r02_mlp_traffic.json)
  cpu_baseline  the reference's own CPU render path on this box's host cores, bounded sample: the UNMODIFIED reference
                staged under oracle/_ref (kind "reference"), or the oracle port if that directory is missing (kind "port")
  e2e           the same metric through the public nerfail_b200.render() with host inputs (pose, intrinsics) and
                the result images copied back to pinned host memory inside the timed region
  strong        the two strong-scaling workloads with a real exchange: attack iteration of 100 views (config 3) and
                retraining step of 4096 rays (config 5), ms per iteration at this N (also under config.strong_scaling_ms)
  extra         the other BASELINE configs in detail (attack iteration, retraining step, 8-NN, view sweep, loader, ...)
--impl reference times the reference's CPU implementation (oracle/_ref: the unmodified run_nerf.render()) on the same
workload, bounded sample per step; rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

H = W = 800
N_SAMPLES, N_IMPORTANCE = 64, 128
FLOP_PER_SAMPLE = 1_186_816          # 593 408 MAC per network evaluation
CHUNK = 1024


def measured_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0)), d.get("hbm_gbs", 6650.0), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        top = sorted(sm)[len(sm) // 2:] if sm else []          # median of the loaded half
        return {"sm_mhz": float(np.median(top)) if top else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class LegoArgs:
    multires, multires_views, i_embed = 10, 4, 0
    use_viewdirs, N_importance, N_samples = True, N_IMPORTANCE, N_SAMPLES
    netdepth = netdepth_fine = 8
    netwidth = netwidth_fine = 256
    netchunk, lrate, perturb, white_bkgd, raw_noise_std = 1 << 16, 5e-4, 1.0, True, 0.0
    dataset_type, no_ndc, lindisp = "blender", False, False
    basedir = expname = ft_path = None
    no_reload = True


def cpu_reference_rays_per_s(n_rays: int, seed_c=0, seed_f=1):
    """The reference's own CPU implementation of the render path (coarse + fine, chunk 1024) on the host cores.
    When oracle/_ref holds the staged, UNMODIFIED reference (oracle/make_ref.py) its run_nerf.render() is what runs —
    the reference's public entry point with rays= , its own NeRF modules, embedders and network_query_fn as create_nerf
    builds them (run_nerf.py:181-204) — kind "reference"; otherwise the oracle port, kind "port".
    Returns (rays/s, seconds, threads, kind)."""
    from oracle import make_ref
    from oracle import nerf_oracle as no
    from oracle import synth
    torch.set_num_threads(os.cpu_count() or 1)
    sd_c = synth.make_non_degenerate(synth.random_state_dict(seed_c), seed_c)
    sd_f = synth.make_non_degenerate(synth.random_state_dict(seed_f), seed_f)
    K, _ = synth.intrinsics(H, W)
    rays = no.camera_rays(H, W, K, torch.tensor(synth.pose_spherical(30.0, -30.0, 4.0)[:3, :4]), 2.0, 6.0)
    g = torch.Generator().manual_seed(0)
    rays = rays[torch.randperm(rays.shape[0], generator=g)[:n_rays]]
    if make_ref.available():
        run_nerf, helpers = make_ref.ref_modules()
        nets = []
        for sd in (sd_c, sd_f):
            n = helpers.NeRF(D=8, W=256, input_ch=63, output_ch=5, skips=[4], input_ch_views=27, use_viewdirs=True)
            n.load_state_dict(sd)
            nets.append(n)
        embed_fn, _ = helpers.get_embedder(10, 0)
        embeddirs_fn, _ = helpers.get_embedder(4, 0)
        query = lambda inputs, viewdirs, network_fn: run_nerf.run_network(inputs, viewdirs, network_fn, embed_fn=embed_fn,
                                                                          embeddirs_fn=embeddirs_fn, netchunk=1024 * 64)
        kw = dict(network_query_fn=query, perturb=False, N_importance=N_IMPORTANCE, network_fine=nets[1], N_samples=N_SAMPLES,
                  network_fn=nets[0], use_viewdirs=True, white_bkgd=True, raw_noise_std=0., ndc=False, lindisp=False,
                  near=2.0, far=6.0)
        batch = (rays[:, 0:3].contiguous(), rays[:, 3:6].contiguous())
        t0 = time.perf_counter()
        with torch.no_grad():
            run_nerf.render(H, W, K, chunk=CHUNK, rays=batch, **kw)
        dt = time.perf_counter() - t0
        return n_rays / dt, dt, torch.get_num_threads(), "reference"
    t0 = time.perf_counter()
    with torch.no_grad():
        for i in range(0, n_rays, CHUNK):
            no.render_ray_batch(rays[i:i + CHUNK], sd_c, sd_f, N_SAMPLES, N_IMPORTANCE, True)
    dt = time.perf_counter() - t0
    return n_rays / dt, dt, torch.get_num_threads(), "port"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = 8192                                   # ~3.5 s of 16-core CPU work per step: K=5, W=3 stays under a minute
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_reference_rays_per_s(256)
    vals, cores, kind = [], 1, "port"
    t_total = 0.0
    for _ in range(args.steps):
        v, dt, cores, kind = cpu_reference_rays_per_s(n)
        vals.append(v); t_total += dt
    value = n * args.steps / t_total
    how = ("the unmodified reference's run_nerf.render() staged under oracle/_ref (oracle/make_ref.py)" if kind == "reference"
           else "oracle/nerf_oracle.py (oracle/_ref absent)")
    line = {
        "impl": "reference", "metric": "render rays/s", "value": value, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "single 800x800 synthetic Blender view render (coarse+fine, 1024-ray chunks)",
                   "N_samples": N_SAMPLES, "N_importance": N_IMPORTANCE, "netwidth": 256, "multires": [10, 4],
                   "sample": f"{n} random rays of the view per step"},
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": cores, "kind": kind,
                         "sample": f"{n} rays x {args.steps} steps, torch CPU fp32, {how}"},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_extras(args, dev, rank, world, dist, nb, ops, synth, kw):
    """Side measurements for the other BASELINE configs (reported under "extra", not part of `value`):
    config 3 — NeRFail attack iteration: GaussNet gather forward + scatter backward over 100 views (800x800, P=3,
               upstream gradient ~ N(0,1), classifier excluded), views sharded over ranks, ONE all-reduce
               of grad_spatial_rgb per iteration (30.72 MB);
    config 5 — NeRF retraining step: 4096 rays forward + backward through coarse + fine (fp32 layer kernels), rays
               sharded over ranks, one all-reduce per network of the parameter gradients."""
    from nerfail_b200 import dist as nd
    out = {}
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def sync_max(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t)
        return ms

    # ---- config 3: attack iteration (gather + scatter over this rank's pixels, exchange + sign step) ----
    P, Hh, Ww, V = 3, 800, 800, 100
    HW = Hh * Ww
    pieces = nd.shard_pixels(V, HW, rank, world, quantum=Ww)          # 100 views over 8 ranks: 12.5 views each, cut on rows
    n_px = sum(e_ - b_ for _, b_, e_ in pieces)
    g = torch.Generator(device="cpu").manual_seed(1234)
    table0 = torch.randn(P, Hh, Ww, 4, generator=g) * 5.0
    yy, xx = torch.meshgrid(torch.arange(Hh, dtype=torch.float32), torch.arange(Ww, dtype=torch.float32), indexing="ij")
    disc = ((yy - Hh / 2) ** 2 + (xx - Ww / 2) ** 2) <= 0.4 * Hh * Ww / np.pi          # A = 255 on a centred disc of ~40 % of the pixels
    table0[..., 3] = disc.float() * 255.0
    T = P * HW
    ex = nd.PeerExchange(T * 4, T * 4, dev)                        # perturbation table + its gradient in NVLink peer memory
    table = ex.value[:T * 4].view(P, Hh, Ww, 4)
    table.copy_(table0.to(dev))
    init = table.clone()
    grad = ex.grad[:T * 4].view(P, Hh, Ww, 4)
    active_fraction = float((table0[..., 3] > 0).float().mean())
    # This rank's pixels of the 100-view batch as ONE contiguous [1,2,n_px,8] weight / index tensor (the GaussNet kernels are
    # per pixel: a shard is a pixel range, dist.shard_pixels), every view with its own random neighbours, generated ten
    # views at a time.  One gather launch and one scatter launch per iteration cover the whole shard.
    gd = torch.Generator(device=dev).manual_seed(100 + rank)
    rows = n_px // Ww
    w_idx = torch.empty((1, 2, rows, Ww, 8), dtype=torch.float32, device=dev)
    ori = torch.empty((1, rows, Ww, 4), dtype=torch.uint8, device=dev)
    gx = torch.empty((1, rows, Ww, 4), dtype=torch.float32, device=dev)
    r0 = 0
    while r0 < rows:
        nr = min(10 * Hh, rows - r0)
        pix = (torch.arange(r0 * Ww, (r0 + nr) * Ww, device=dev) % HW).reshape(1, nr, Ww, 1)      # pixel index inside its view
        view_of_row = torch.randint(0, P, ((nr + Hh - 1) // Hh + 1, 1, 1), device=dev, generator=gd).repeat_interleave(Hh, 0)[:nr]
        idx = (pix + view_of_row.reshape(1, nr, 1, 1) * HW
               + torch.randint(-400, 401, (1, nr, Ww, 8), device=dev, generator=gd)).clamp_(0, T - 1).float()
        dist_ = torch.sort(torch.randn(1, nr, Ww, 8, device=dev, generator=gd).abs() * 0.01, dim=-1).values
        w_idx[:, :, r0:r0 + nr] = ops.gauss_weights(torch.stack([dist_, idx], 1), 0.02)
        ori[:, r0:r0 + nr] = torch.randint(0, 256, (1, nr, Ww, 4), device=dev, generator=gd, dtype=torch.uint8)
        gx[:, r0:r0 + nr] = torch.randn(1, nr, Ww, 4, device=dev, generator=gd)
        del idx, dist_, pix
        r0 += nr
    tshape = (P, Hh, Ww, 4)
    side = torch.cuda.Stream(device=dev)
    grad_ready = torch.cuda.Event()

    def scatter_views(overlap_zero=False):
        main = torch.cuda.current_stream(dev)
        if overlap_zero:            # the gradient is zeroed on a side stream beside the gather (which does not touch it)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                grad.zero_()
                grad_ready.record(side)
        else:
            grad.zero_()
        x, x_rgba = ops.gauss_gather_fwd(table.reshape(-1, 4), w_idx, ori, 32.0)
        if overlap_zero:
            main.wait_event(grad_ready)
        ops.gauss_scatter_bwd(None, gx, x, w_idx, ori, 32.0, tshape, out=grad)

    def attack_iter():
        # attack_NeRFail_S.py:317-392 minus the classifier: zero the gradient, forward + backward of this rank's pixels,
        # then ONE kernel per GPU that sums the partial gradients over NVLink peer memory, takes the sign step on the rows it
        # owns and writes them into every rank's table (csrc/peer.cu) - the I-FGSM update is inside the timed iteration
        scatter_views(overlap_zero=True)
        ex.attack_step(init.reshape(-1, 4), T, 1.0, 8.0)

    def attack_iter_nccl():
        # the round-1 form for comparison: NCCL all-reduce of the whole gradient table, then the update as its own kernels
        scatter_views()
        nd.attack_sign_step_(table, grad, init, 1.0, 8.0)

    def time_iters(fn, iters=6):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return sync_max(e0.elapsed_time(e1)) / iters

    ms = time_iters(attack_iter)
    ex.status()
    ms_nccl = time_iters(attack_iter_nccl)
    ms_local = time_iters(lambda: scatter_views(overlap_zero=True))
    px = V * HW
    slice_rows = (T + world - 1) // world
    nvlink_bytes = int(active_fraction * slice_rows * 16 * (world - 1))
    out["attack_iteration"] = {"metric": "GaussNet fwd+bwd attack rays/s (1 pixel = 1 ray)", "value": px / (ms / 1e3), "unit": "rays/s",
                               "ms_per_iteration": ms, "views": V, "pixels_per_rank": n_px, "views_per_rank": n_px / HW,
                               "exchange": "one kernel per GPU over NVLink peer memory: P2P reduce of the owned rows + sign step + P2P broadcast (csrc/peer.cu)",
                               "sign_step_in_timed_region": True,
                               "nvlink_bytes_read_per_gpu": nvlink_bytes, "nvlink_bytes_written_per_gpu": nvlink_bytes,
                               "ms_gather_scatter_only": ms_local, "ms_exchange_and_update": max(0.0, ms - ms_local),
                               "note_exchange": "difference of two separately timed loops (run-to-run spread ~2 %)",
                               "ms_per_iteration_nccl_allreduce_then_update": ms_nccl, "nccl_allreduce_bytes": T * 16,
                               "active_fraction": active_fraction,
                               "algorithmic_bytes_per_pixel": 456, "achieved_GBps_per_gpu": n_px * 456 / (ms / 1e3) / 1e9,
                               "scaling": "strong"}
    del w_idx, ori, gx
    ex.close()
    del table, grad, init, ex

    # ---- config 5: retraining step ----
    N_rand = 4096
    b, e = nd.shard_range(N_rand, rank, world)
    K, _ = synth.intrinsics(H, W)
    rays_all = ops.get_ray_batch(H, W, K, torch.tensor(synth.camera_ring(8)[1][:3, :4]), 2.0, 6.0, device=dev)
    sel = torch.from_numpy(np.random.default_rng(0).choice(H * W, N_rand, replace=False)).to(dev)
    target_all = torch.rand(N_rand, 3, generator=torch.Generator().manual_seed(5))
    batch = {"rays": rays_all[sel][b:e].contiguous(), "target": target_all[b:e].to(dev)}       # strong: the batch is split over ranks
    sel_w = torch.from_numpy(np.random.default_rng(100 + rank).choice(H * W, N_rand, replace=False)).to(dev)
    batch_weak = {"rays": rays_all[sel_w].contiguous(), "target": target_all.to(dev)}            # weak: every rank brings N_rand rays
    kwt = {k: v for k, v in kw.items() if k not in ("use_viewdirs", "ndc")}
    kwt.update(perturb=1.0)
    params_c, params_f = list(kw["network_fn"].parameters()), list(kw["network_fine"].parameters())

    def train_step():
        for p_ in params_c + params_f:
            p_.grad = None
        with torch.enable_grad():
            ret = nb.render_rays(batch["rays"], retraw=True, **kwt)
            loss = nb.img2mse(ret["rgb_map"], batch["target"]) + nb.img2mse(ret["rgb0"], batch["target"])
            loss.backward()
        nd.allreduce_grads_(params_c, scale=1.0 / world)
        nd.allreduce_grads_(params_f, scale=1.0 / world)
        return loss

    def time_train(n_iter):
        nonlocal train_step
        train_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(n_iter):
            train_step()
        e1.record()
        torch.cuda.synchronize()
        return sync_max(e0.elapsed_time(e1)) / n_iter

    ar_bytes = int(sum(p_.numel() for p_ in params_c + params_f) * 4)
    prev = os.environ.get("NERFAIL_B200_TRAIN")
    os.environ["NERFAIL_B200_TRAIN"] = "fp32"       # exact-parity layer kernels (reported next to the tensor-core step)
    ms32 = time_train(2)
    opt = nb.Adam(params_c + params_f, lr=5e-4, betas=(0.9, 0.999))      # run_nerf.py:213; fused multi-tensor kernel
    plain_step = train_step

    adam_steps = [0]

    def train_step_adam():                                               # run_nerf.py:776-800: backward, step, lr decay
        loss = plain_step()
        opt.step()
        adam_steps[0] += 1                                               # host-side global_step: no device sync in the loop
        nb.set_lrate(opt, nb.decayed_lrate(5e-4, 250, adam_steps[0]))
        return loss
    os.environ["NERFAIL_B200_TRAIN"] = "bf16"       # fused tensor-core forward (saves activations) + dgrad chain + wgrad GEMMs
    try:
        ms16 = time_train(5)
        sd0 = [p_.detach().clone() for p_ in params_c + params_f]
        train_step = train_step_adam
        ms16_adam = time_train(5)
        strong_batch, batch = batch, batch_weak
        ms16_adam_weak = time_train(5)
        batch = strong_batch
        train_step = plain_step

        # the same step (render -> loss -> backward -> all-reduce -> fused Adam -> lr decay -> re-pack) as ONE CUDA graph
        from nerfail_b200 import train as ntrain

        def time_graphed(bt, n_iter, peer=False):
            opt_g = nb.Adam(params_c + params_f, lr=5e-4, betas=(0.9, 0.999))
            # peer: gradient average + Adam (sharded state) + parameter broadcast as ONE kernel per GPU over NVLink peer
            # memory (dist.PeerAdam, csrc/peer.cu) instead of two NCCL all-reduces and a replicated full-size Adam
            pex = nd.PeerAdam([net_c_, net_f_], opt_g, dev) if peer else None
            br = torch.stack([bt["rays"][:, 0:3], bt["rays"][:, 3:6]], 0).contiguous()
            kwg = dict(kw, perturb=1.0)
            stepper = ntrain.GraphedTrainStep(br.shape[1], H, W, K, 32768, kwg, opt_g, 5e-4, 250, near=2.0, far=6.0, device=dev,
                                              exchange=pex)
            for i in range(5):                                            # 3 eager warm-up steps, capture, one replay
                stepper(br, bt["target"], i)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = ev(), ev()
            e0.record()
            for i in range(n_iter):
                stepper(br, bt["target"], 5 + i)
            e1.record()
            torch.cuda.synchronize()
            t = sync_max(e0.elapsed_time(e1)) / n_iter
            if pex is not None:
                pex.ex.status()
                del stepper
                pex.close()
            return t
        net_c_, net_f_ = kw["network_fn"], kw["network_fine"]
        ms_graph = time_graphed(batch, 20)
        ms_graph_weak = time_graphed(batch_weak, 20)
        ms_graph_peer = time_graphed(batch, 20, peer=True)
        ms_graph_peer_weak = time_graphed(batch_weak, 20, peer=True)
        with torch.no_grad():                                            # restore the weights the other measurements use
            for p_, w_ in zip(params_c + params_f, sd0):
                p_.copy_(w_)
    finally:
        if prev is None:
            os.environ.pop("NERFAIL_B200_TRAIN", None)
        else:
            os.environ["NERFAIL_B200_TRAIN"] = prev
    flop_step = N_rand * (N_SAMPLES + N_SAMPLES + N_IMPORTANCE) * 3_489_024        # fwd + dgrad + wgrad
    out["retraining_step"] = {"metric": "NeRF fwd+bwd rays/s (4096-ray batch, coarse+fine, bf16 tensor-core training kernels)",
                              "value": N_rand / (ms16 / 1e3), "unit": "rays/s", "ms_per_step": ms16, "rays_per_rank": int(e - b),
                              "dtype": "bf16", "optimizer_step": False, "allreduce_bytes": ar_bytes, "scaling": "strong",
                              "with_fused_adam_and_weight_repack": {"ms_per_step": ms16_adam, "value": N_rand / (ms16_adam / 1e3), "unit": "rays/s"},
                              "cuda_graph_step_with_adam": {"ms_per_step": ms_graph, "value": N_rand / (ms_graph / 1e3), "unit": "rays/s",
                                                            "note": "whole optimisation step captured once (train.GraphedTrainStep), 20 replays; NCCL all-reduce + replicated Adam"},
                              "cuda_graph_step_peer_adam": {"ms_per_step": ms_graph_peer, "value": N_rand / (ms_graph_peer / 1e3), "unit": "rays/s",
                                                            "note": "same graph with dist.PeerAdam: average + sharded Adam + broadcast in one kernel per GPU over NVLink peer memory"},
                              "cuda_graph_step_peer_adam_weak": {"ms_per_step": ms_graph_peer_weak, "rays_per_rank": N_rand, "global_batch": N_rand * world,
                                                                 "value": N_rand * world / (ms_graph_peer_weak / 1e3), "unit": "rays/s"},
                              "cuda_graph_step_with_adam_weak": {"ms_per_step": ms_graph_weak, "rays_per_rank": N_rand, "global_batch": N_rand * world,
                                                                 "value": N_rand * world / (ms_graph_weak / 1e3), "unit": "rays/s"},
                              "weak_scaling_with_adam": {"ms_per_step": ms16_adam_weak, "rays_per_rank": N_rand, "global_batch": N_rand * world,
                                                         "value": N_rand * world / (ms16_adam_weak / 1e3), "unit": "rays/s"},
                              "algorithmic_TFLOPs": flop_step / world / (ms16 / 1e3) / 1e12,
                              "fp32_layer_kernels": {"value": N_rand / (ms32 / 1e3), "unit": "rays/s", "ms_per_step": ms32, "dtype": "f32"}}
    for p_ in params_c + params_f:
        p_.grad = None

    # ---- 8-NN precompute (create_index_and_dist.py:110-163) on rendered geometry: one view, and the sweep sharded by view ----
    from nerfail_b200 import pipeline
    poses = synth.camera_ring(8)
    kwr = {k: v for k, v in kw.items()}
    kwr.update(near=2.0, far=6.0)
    base = torch.stack([pipeline.render_points(H, W, K, torch.tensor(poses[i][:3, :4]), 1024, **kwr) for i in (0, 3, 5)], 0)
    qpts = pipeline.render_points(H, W, K, torch.tensor(poses[1 + rank % 2][:3, :4]), 1024, **kwr)
    cand = base.reshape(-1, 3).contiguous()
    # every rank answers 8 views of the sweep (weak scaling: view i -> rank i mod G, candidates replicated, no collective),
    # index_and_dist + index_and_weight per view as the reference stores them; the views are this rank's rendered points
    # shifted by a few 1e-4 so that no two queries are identical
    sps = pipeline.SpatialPointSet(base)
    n_knn = 8
    my_views = {i: qpts + 1e-4 * (i + 1) for i in range(rank, n_knn * world, world)}
    view_list = [my_views.get(i) for i in range(n_knn * world)]
    pipeline.knn_sweep(view_list[:world], sps, rank=rank, world_size=world, keep=False)      # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = ev(), ev()
    e0.record()
    pipeline.knn_sweep(view_list, sps, rank=rank, world_size=world, keep=False)
    e1.record()
    torch.cuda.synchronize()
    ms_knn = sync_max(e0.elapsed_time(e1))
    out["knn_sweep"] = {"metric": "8-NN precompute sweep, views/s (800x800 query views against P=3 base views, exact grid search + Gaussian weights)",
                        "value": n_knn * world / (ms_knn / 1e3), "unit": "views/s", "views": n_knn * world, "ms_per_view_per_rank": ms_knn / n_knn,
                        "query_pixels_per_s": n_knn * world * H * W / (ms_knn / 1e3), "scaling": "weak",
                        "sharding": "view i -> rank i mod G, candidates replicated (23 MB), no collective"}
    del my_views, view_list, sps
    if world == 1:

        def timed(fn, n):
            fn(); torch.cuda.synchronize()
            a, b_ = ev(), ev()
            a.record()
            for _ in range(n):
                r = fn()
            b_.record(); torch.cuda.synchronize()
            return a.elapsed_time(b_) / n, r

        ms_build, grid = timed(lambda: ops.KnnGrid(cand), 2)
        stats = torch.zeros(1, dtype=torch.int64, device=dev)
        ms_grid, (d_g, i_g) = timed(lambda: grid.query(qpts.reshape(-1, 3), stats), 3)
        evals = float(stats.item()) / 4.0
        ms_brute, (d_b, i_b) = timed(lambda: ops.knn8(qpts.reshape(-1, 3), cand), 1)
        pairs = float(qpts.numel() // 3) * float(cand.shape[0])
        out["knn_view"] = {"metric": "exact 8-NN of one 800x800 view against P=3 base views (1.92 M candidates)",
                           "grid_ms": ms_grid, "grid_build_ms": ms_build, "brute_force_ms": ms_brute,
                           "identical_to_brute_force": bool(torch.equal(i_g, i_b) and torch.equal(d_g, d_b)),
                           "pruning_factor": pairs / max(evals, 1.0), "brute_force_pairs_per_s": pairs / (ms_brute / 1e3),
                           "speedup": ms_brute / ms_grid}

    # ---- novel-view sweep through render_path (nerf_render_only.py:619-648): PNG + host arrays, asynchronous sink ----
    import shutil
    import tempfile
    n_sweep = 12
    ring = [torch.tensor(p_, dtype=torch.float32) for p_ in synth.camera_ring(n_sweep * world)]
    tmp = tempfile.mkdtemp(prefix="nfb_sweep_")
    try:
        kws = {k: v for k, v in kw.items()}
        kws.update(near=2.0, far=6.0)
        with torch.no_grad():
            nb.render_sweep(ring[:world], (H, W, K[0][0]), K, 1024, kws, savedir=tmp, rank=rank, world_size=world)   # warm-up
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            nb.render_sweep(ring, (H, W, K[0][0]), K, 1024, kws, savedir=tmp, rank=rank, world_size=world)
            torch.cuda.synchronize()
            ms_sweep = sync_max(1e3 * (time.perf_counter() - t0))
        out["view_sweep"] = {"metric": "novel-view sweep rays/s through render_sweep (views sharded i mod G; rgb/disp to host, PNG files written)",
                             "value": n_sweep * world * H * W / (ms_sweep / 1e3), "unit": "rays/s", "views": n_sweep * world,
                             "ms_per_view_per_rank": ms_sweep / n_sweep, "scaling": "weak", "timing": "host wall clock incl. file writes, max over ranks"}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)

    # ---- Blender loader (load_blender.py:37-110): 48 RGBA PNGs of 800x800, thread-pool decode vs one after the other ----
    if rank == 0:
        from nerfail_b200 import data as ndata
        tmp = tempfile.mkdtemp(prefix="nfb_scene_")
        try:
            synth.write_blender_scene(tmp, 800, 800, (16, 16, 16), seed=0)
            t0 = time.perf_counter()
            imgs, poses, _, hwf, _ = ndata.load_blender_data(tmp, half_res=False, testskip=1)
            t_pool = time.perf_counter() - t0
            t0 = time.perf_counter()
            ndata.load_blender_data(tmp, half_res=False, testskip=1, workers=1)
            t_serial = time.perf_counter() - t0
            out["blender_loader"] = {"metric": "load_blender_data, 48 RGBA PNGs of 800x800 (incompressible noise images: worst case for the decoder)",
                                     "images_per_s": imgs.shape[0] / t_pool, "seconds": t_pool, "seconds_one_thread": t_serial,
                                     "threads": min(32, os.cpu_count() or 1)}
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the attack-iteration / retraining-step side measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=dev)

    import nerfail_b200 as nb
    from nerfail_b200 import _lib, ops
    from oracle import synth          # seeded synthetic weights / cameras only (generators, no arithmetic under test)

    _, kw, *_ = nb.create_nerf(LegoArgs(), device=dev)
    kw["network_fn"].load_state_dict(synth.make_non_degenerate(synth.random_state_dict(0), 0))
    kw["network_fine"].load_state_dict(synth.make_non_degenerate(synth.random_state_dict(1), 1))
    net_c, net_f = kw["network_fn"], kw["network_fine"]
    K, _ = synth.intrinsics(H, W)
    poses = synth.camera_ring(max(8, world))
    c2w_host = torch.tensor(poses[rank % len(poses)][:3, :4])
    n_rays = H * W
    from nerfail_b200.rendering import MAX_RAYS_PER_PASS as rays_per_pass

    # ---------------- device-resident step (value) ----------------
    rays = ops.get_ray_batch(H, W, K, c2w_host, 2.0, 6.0, device=dev)
    fc, ff = net_c.fused(), net_f.fused()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    mlp_events = []

    def step_resident(record=False):
        outs = []
        for i in range(0, n_rays, rays_per_pass):
            rb = rays[i:i + rays_per_pass]
            z = ops.coarse_z(rb, N_SAMPLES)
            e = [ev() for _ in range(4)] if record else None
            if record: e[0].record()
            raw = fc.forward_rays(rb, z)
            if record: e[1].record()
            _, _, _, wts, _ = ops.composite_fwd(raw, z, rb, None, True)
            zf, _, zstd = ops.hierarchical(z, wts, N_IMPORTANCE)
            if record: e[2].record()
            raw = ff.forward_rays(rb, zf)
            if record: e[3].record()
            outs.append(ops.composite_fwd(raw, zf, rb, None, True))
            if record:
                mlp_events.append((e, rb.shape[0]))
        return outs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(args.warmup):
            step_resident()
        fc.status(); ff.status()
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        l0 = _lib.launch_count()
        t_start, t_stop = ev(), ev()
        t_start.record()
        for _ in range(args.steps):
            step_resident(record=True)
        t_stop.record()
        barrier()
        launches = _lib.launch_count() - l0
        ms = t_start.elapsed_time(t_stop)
        clocks = sampler.stop() if rank == 0 else None
        fc.status(); ff.status()

        # ---------------- end-to-end step through the public API (e2e) ----------------
        pin = {k: torch.empty(s, dtype=torch.float32).pin_memory() for k, s in
               (("rgb", (H, W, 3)), ("disp", (H, W)), ("acc", (H, W)))}

        def step_e2e():
            rgb, disp, acc, _ = nb.render(H, W, K, chunk=CHUNK, c2w=c2w_host, near=2.0, far=6.0, **kw)   # host pose in
            pin["rgb"].copy_(rgb, non_blocking=True); pin["disp"].copy_(disp, non_blocking=True)
            pin["acc"].copy_(acc, non_blocking=True)
            torch.cuda.current_stream().synchronize()                                                    # images on the host

        for _ in range(2):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        e_start, e_stop = ev(), ev()
        e_start.record()
        for _ in range(args.steps):
            step_e2e()
        e_stop.record()
        barrier()
        ms_e2e = max(e_start.elapsed_time(e_stop), 1e3 * (time.perf_counter() - t0))

    # the same view with the reference's LITERAL chunking (625 passes of 1024 rays, NERFAIL_B200_STRICT_CHUNK=1): what
    # chunk=1024 costs when it is not coalesced (results are bit-identical, tests/test_gpu_render.py)
    ms_strict = None
    if not args.no_extras:
        import nerfail_b200.rendering  # noqa: F401
        os.environ["NERFAIL_B200_STRICT_CHUNK"] = "1"
        try:
            with torch.no_grad():
                nb.render(H, W, K, chunk=CHUNK, c2w=c2w_host, near=2.0, far=6.0, **kw)
                barrier()
                s0, s1 = ev(), ev()
                s0.record()
                nb.render(H, W, K, chunk=CHUNK, c2w=c2w_host, near=2.0, far=6.0, **kw)
                s1.record()
                barrier()
                ms_strict = s0.elapsed_time(s1)
        finally:
            os.environ.pop("NERFAIL_B200_STRICT_CHUNK", None)

    extra = {}
    if not args.no_extras:
        extra = run_extras(args, dev, rank, world, dist, nb, ops, synth, kw)

    # MLP kernel time (both launches of every pass), measured inside the timed region on the launching stream
    mlp_ms = sum(e[0].elapsed_time(e[1]) + e[2].elapsed_time(e[3]) for e, _ in mlp_events)
    mlp_flop = sum(r * (N_SAMPLES + N_SAMPLES + N_IMPORTANCE) * FLOP_PER_SAMPLE for _, r in mlp_events)
    n_mlp_launches = 2 * len(mlp_events)

    if world > 1:
        t = torch.tensor([ms, ms_e2e, mlp_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, mlp_ms = (float(x) for x in t)
    total_rays = n_rays * args.steps * world
    value = total_rays / (ms / 1e3)
    e2e_value = total_rays / (ms_e2e / 1e3)

    if rank == 0:
        peak_tf, _, how = measured_peaks()
        achieved_tf = mlp_flop / (mlp_ms / 1e3) / 1e12
        # dram__bytes_read.sum + dram__bytes_write.sum of the committed `ncu --set full` capture of this kernel
        # (profiles/r02_mlp_traffic.json, written by scripts/ncu_traffic.py from the raw CSV export next to it), per sample,
        # scaled to this run's average launch; null when no capture is committed
        traffic, traffic_unit = None, "no committed ncu capture (profiles/r02_mlp_traffic.json missing)"
        tpath = os.path.join(REPO, "profiles", "r02_mlp_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            traffic = tj["dram_bytes_per_sample"] * (mlp_flop / FLOP_PER_SAMPLE) / max(1, n_mlp_launches)
            traffic_unit = (f"bytes per launch = {tj['dram_bytes_per_sample']:.2f} B/sample (dram__bytes_read.sum + dram__bytes_write.sum of "
                            f"{tj['source']}) x samples per launch; algorithmic 20 B/sample")
        line = {
            "metric": "render rays/s", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "single 800x800 synthetic Blender view render (coarse+fine, 1024-ray chunks)",
                       "views_per_step": world, "rays_per_step": n_rays * world, "N_samples": N_SAMPLES,
                       "N_importance": N_IMPORTANCE, "netwidth": 256, "netdepth": 8, "multires": [10, 4],
                       "weights": "random-init, sigma head affine-normalised (SURVEY 8d W-B)", "chunk": CHUNK,
                       "chunk_coalescing": f"render() merges consecutive chunks up to {rays_per_pass} rays per kernel pass "
                                           "(chunk does not affect results, run_nerf.py:78-79)",
                       "l2": "per-step working set (1.97 GB of raw network outputs + 0.5 GB depths/weights) exceeds the 126 MB L2; no flush",
                       "parallelism": f"view-sharded x{world}" if world > 1 else "single GPU"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": 12 * 4 + 9 * 8,
                    "d2h_bytes_per_step": H * W * 5 * 4, "ms_per_step": ms_e2e / args.steps,
                    "api": "nerfail_b200.render(H, W, K, chunk=1024, c2w=<host pose>) + rgb/disp/acc to pinned host memory"},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "nfb::mlp_fused_fwd_kernel", "bound": "tensor", "achieved": achieved_tf, "peak": peak_tf,
                         "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                         "traffic": traffic, "traffic_unit": traffic_unit,
                         "peak_source": f"{how} bf16_tflops_sustained", "launches": n_mlp_launches,
                         "avg_launch_ms": mlp_ms / max(1, n_mlp_launches),
                         "algorithmic_flop_per_sample": FLOP_PER_SAMPLE, "share_of_step": mlp_ms / ms},
        }
        if extra:
            # the two workloads with a real exchange, strong scaling (fixed total work), at the top level so that the record
            # of every N carries them: attack iteration of 100 views, retraining step of 4096 rays (one CUDA graph, PeerAdam)
            line["strong"] = {"attack_ms": extra["attack_iteration"]["ms_per_iteration"],
                              "train_ms": extra["retraining_step"]["cuda_graph_step_peer_adam"]["ms_per_step"],
                              "attack_ms_nccl": extra["attack_iteration"]["ms_per_iteration_nccl_allreduce_then_update"],
                              "train_ms_nccl": extra["retraining_step"]["cuda_graph_step_with_adam"]["ms_per_step"],
                              "knn_views_per_s": extra["knn_sweep"]["value"]}
            # the same two figures inside a key every record keeps whole
            line["config"]["strong_scaling_ms"] = {"attack_iteration_100_views": round(line["strong"]["attack_ms"], 4),
                                                   "retraining_step_4096_rays": round(line["strong"]["train_ms"], 4)}
            if ms_strict is not None:
                extra["render_strict_chunk_1024"] = {"metric": "render rays/s with literal 1024-ray chunks (625 passes per view, NERFAIL_B200_STRICT_CHUNK=1)",
                                                     "value": n_rays / (ms_strict / 1e3), "unit": "rays/s", "ms_per_view": ms_strict,
                                                     "note": "host-bound: ~8 kernel launches per 1024-ray pass through Python"}
            line["extra"] = extra
        if world == 1 and not args.no_cpu_baseline:
            n_cpu = 32768                                         # ~14 s of CPU work on a 16-core host
            v, dt, cores, kind = cpu_reference_rays_per_s(n_cpu)
            line["cpu_baseline"] = {"value": v, "unit": "rays/s", "cores": cores, "kind": kind,
                                    "sample": f"{n_cpu} random rays of the same view, coarse+fine, chunk 1024, {dt:.1f} s of torch-CPU fp32 "
                                              + ("(unmodified reference run_nerf.render(), oracle/_ref)" if kind == "reference" else "(oracle/nerf_oracle.py)")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
