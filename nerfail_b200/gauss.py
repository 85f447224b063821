"""Host-side mirror of model/GaussNet.py (gauss_net, create_gauss_w) and of the 8-NN precompute.

Same constructor / forward signatures and return tuples as the reference; the gather, the Gaussian
weighting, the epsilon clip / composite and their backward run in libnerfail_b200.so.
"""
from __future__ import annotations

import os

import torch
from torch import nn

from . import ops


class create_gauss_w(nn.Module):
    """model/GaussNet.py:161-186: distances -> normalised Gaussian weights; returns (i_w, dist)."""

    def __init__(self, device, c):
        super().__init__()
        self.top_number = 8
        self.device = device
        self.c = c

    def forward(self, dist_and_index_list):
        di = dist_and_index_list
        if di.shape[-1] != 8 or di.shape[1] != 2:
            raise ValueError("expected [B,2,H,W,8] (dist, index)")
        i_w = ops.gauss_weights(di, float(self.c))
        dist = di[:, 0, :, :, :].unsqueeze(1)
        return i_w, dist


class gauss_net(nn.Module):
    """model/GaussNet.py:7-159.

    forward(spatial_rgb [P,H,W,4], weight_and_index_list [B,2,H,W,8], ori_img [B,H,W,4] uint8)
      -> (x [B,H,W,4], x_rgba [B,H,W,4], cla, ori_img float, ori_cla)
    x / x_rgba come from one gather kernel; the running epsilon_3d_min/max (reference :89-103, two host syncs
    per call) is tracked on the device and only read back by print_epsilon()/the properties.
    """

    def __init__(self, device, c, model, model_name, epsilon=None):
        super().__init__()
        self.top_number = 8
        self.c = torch.nn.Parameter(torch.tensor([c]), requires_grad=False)
        self.device = device
        self.model = model
        self.model_name = model_name
        self.epsilon = epsilon
        self.update_epsilon_3d = True
        self._minmax = None
        self.w = 299
        self.h = 299
        self.resize_antialias = None     # None: what torchvision.transforms.Resize of the installed torchvision does

    # -- epsilon tracking (device resident) --
    def _mm(self, device):
        if self._minmax is None or self._minmax.device != device:
            self._minmax = torch.zeros(2, dtype=torch.float32, device=device)
        return self._minmax

    @property
    def epsilon_3d_max(self):
        return 0 if self._minmax is None else float(self._minmax[1])

    @property
    def epsilon_3d_min(self):
        return 0 if self._minmax is None else float(self._minmax[0])

    def epsilon_3d_zero(self):
        if self._minmax is not None:
            self._minmax.zero_()

    def close_update_epsilon_3d(self):
        self.update_epsilon_3d = False

    def open_update_epsilon_3d(self):
        self.update_epsilon_3d = True

    def print_epsilon(self):
        print("epsilon_3d_min: ", self.epsilon_3d_min)
        print("epsilon_3d_max: ", self.epsilon_3d_max)

    def perturbed(self, spatial_rgb, weight_and_index_list, ori_img):
        """(x, x_rgba) only — the part of forward() this library accelerates (reference :53-119)."""
        if not spatial_rgb.is_cuda:
            raise RuntimeError("nerfail_b200.gauss_net runs on CUDA only")
        w_idx = weight_and_index_list.float().contiguous()
        ori = ori_img
        if ori.dtype != torch.uint8:
            ori = ori.to(torch.uint8)
        ori = ori.contiguous()
        mm = self._mm(spatial_rgb.device) if self.update_epsilon_3d else None
        if spatial_rgb.requires_grad and torch.is_grad_enabled():
            return ops.GaussGatherFn.apply(spatial_rgb, w_idx, ori, self.epsilon, mm)
        return ops.gauss_gather_fwd(spatial_rgb.detach().float().reshape(-1, 4).contiguous(), w_idx, ori, self.epsilon, mm)

    def classifier_size(self):
        """Side of the classifier's input: None for my_model (no Resize), 224 for vit_b_16, else 299 (reference :147-154)."""
        if self.model_name == "my_model":
            return None
        return 224 if self.model_name == "vit_b_16" else 299

    def _to_classifier_input(self, img_bhwc):
        """RGBA -> NCHW RGB on white (reference :121-145) and, unless my_model, the bilinear Resize of :147-154 — ONE kernel
        (nfb_rgba_to_chw / nfb_rgba_to_chw_resized), differentiable twice.  `resize_antialias` follows the installed
        torchvision's default for tensors (ops.default_resize_antialias)."""
        size = self.classifier_size()
        if size is None:
            if img_bhwc.dtype == torch.uint8:
                return ops.rgba_u8_to_chw(img_bhwc)
            return ops.RgbaToChwFn.apply(img_bhwc, 255.0)
        aa = self.resize_antialias if self.resize_antialias is not None else ops.default_resize_antialias()
        if img_bhwc.dtype == torch.uint8:
            return ops.rgba_u8_to_chw_resized(img_bhwc, size, aa)
        return ops.RgbaToChwResizedFn.apply(img_bhwc, 255.0, size, aa)

    def forward(self, spatial_rgb, weight_and_index_list, ori_img, zero_init_mask: bool = False):
        x, x_rgba = self.perturbed(spatial_rgb, weight_and_index_list, ori_img)
        ori_f = ori_img.float() if isinstance(ori_img, torch.Tensor) else torch.tensor(ori_img, dtype=torch.float)
        cla_x = self._to_classifier_input(x_rgba)
        cla_ori = self._to_classifier_input(ori_img if isinstance(ori_img, torch.Tensor) and ori_img.dtype == torch.uint8 and ori_img.is_cuda
                                            else ori_f)
        cla = self.model(cla_x)
        ori_cla = self.model(cla_ori)
        return x, x_rgba, cla, ori_f, ori_cla

    def class_gradients(self, spatial_rgb, weight_and_index_list, ori_img, classes):
        """d cla[:, k].sum() / d spatial_rgb for every k in `classes`, stacked [len(classes), *spatial_rgb.shape], plus cla.

        What deepfool.py:72-86 obtains with two torch.autograd.grad calls per class (14 per iteration for 8 classes), each
        re-walking the Resize adjoint and an 8-neighbour scatter: here the classifier is differentiated for all classes at
        once (is_grads_batched; a loop over retained graphs if the model has an op without a batching rule), the cotangents
        go through ONE launch of the fused Resize / RGBA adjoint (nfb_chw_resized_to_rgba, NC x B images) and ONE launch of
        the batched scatter (nfb_gauss_scatter_bwd_batched).  The gradients are first-order (DeepFool detaches `dr`,
        deepfool.py:98); use forward() with torch.autograd.grad when a graph through them is needed."""
        classes = [int(k) for k in classes]
        with torch.no_grad():
            x, x_rgba = self.perturbed(spatial_rgb.detach(), weight_and_index_list, ori_img)
        w_idx = weight_and_index_list.float().contiguous()
        ori = (ori_img if ori_img.dtype == torch.uint8 else ori_img.to(torch.uint8)).contiguous()
        size = self.classifier_size()
        with torch.no_grad():
            inp = self._to_classifier_input(x_rgba)
        inp = inp.detach().requires_grad_(True)
        with torch.enable_grad():
            cla = self.model(inp)
            tot = cla.sum(0)[classes]                                     # [NC]
            eye = torch.eye(len(classes), dtype=tot.dtype, device=tot.device)
            try:
                g_in = torch.autograd.grad(tot, inp, grad_outputs=eye, is_grads_batched=True)[0]
            except Exception:
                g_in = torch.stack([torch.autograd.grad(tot[i], inp, retain_graph=True)[0] for i in range(len(classes))], 0)
        NC, B = len(classes), x_rgba.shape[0]
        H, W = x_rgba.shape[1], x_rgba.shape[2]
        g_in = g_in.reshape(NC * B, 3, g_in.shape[-2], g_in.shape[-1]).contiguous()
        if size is None:      # my_model: the adjoint of the plain RGBA -> CHW conversion, cotangent n uses image n % B
            g_img = torch.cat([ops.ChwToRgbaFn.apply(g_in[c * B:(c + 1) * B], x_rgba) for c in range(NC)], 0)
        else:
            aa = self.resize_antialias if self.resize_antialias is not None else ops.default_resize_antialias()
            g_img = ops.chw_resized_to_rgba(g_in, x_rgba, H, W, aa)
        grads = ops.gauss_scatter_bwd_batched(g_img.reshape(NC, B, H, W, 4), x, w_idx, ori, self.epsilon, tuple(spatial_rgb.shape))
        return grads, cla.detach()


def knn_index_and_dist(query_hw3: torch.Tensor, base_points) -> torch.Tensor:
    """One view of create_index_and_dist.py:110-151: float32 [2,H,W,8] = cat([dist, idx]) for every pixel's
    3-D point against the [P*H*W,3] base-view points.  `base_points` is the point tensor or, to amortise the grid
    build over the views of a data set, an `ops.KnnGrid` made from it; NERFAIL_B200_KNN=brute selects the plain scan."""
    if isinstance(base_points, ops.KnnGrid):
        return base_points.query_dist_idx(query_hw3)
    if os.environ.get("NERFAIL_B200_KNN", "grid").lower() == "brute":
        return ops.knn8_dist_idx(query_hw3, base_points)
    return ops.KnnGrid(base_points).query_dist_idx(query_hw3)
