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

    @staticmethod
    def _to_classifier_input(img_bhwc):
        """RGBA -> NCHW RGB on white (reference :121-145): one kernel (nfb_rgba_to_chw), differentiable twice."""
        if img_bhwc.dtype == torch.uint8:
            return ops.rgba_u8_to_chw(img_bhwc)
        return ops.RgbaToChwFn.apply(img_bhwc, 255.0)

    def forward(self, spatial_rgb, weight_and_index_list, ori_img, zero_init_mask: bool = False):
        x, x_rgba = self.perturbed(spatial_rgb, weight_and_index_list, ori_img)
        ori_f = ori_img.float() if isinstance(ori_img, torch.Tensor) else torch.tensor(ori_img, dtype=torch.float)
        cla_x = self._to_classifier_input(x_rgba)
        cla_ori = self._to_classifier_input(ori_img if isinstance(ori_img, torch.Tensor) and ori_img.dtype == torch.uint8 and ori_img.is_cuda
                                            else ori_f)
        if self.model_name != "my_model":
            size = 224 if self.model_name == "vit_b_16" else 299
            from torchvision.transforms import Resize
            rs = Resize([size, size])
            cla_x, cla_ori = rs(cla_x), rs(cla_ori)
        cla = self.model(cla_x)
        ori_cla = self.model(cla_ori)
        return x, x_rgba, cla, ori_f, ori_cla


def knn_index_and_dist(query_hw3: torch.Tensor, base_points) -> torch.Tensor:
    """One view of create_index_and_dist.py:110-151: float32 [2,H,W,8] = cat([dist, idx]) for every pixel's
    3-D point against the [P*H*W,3] base-view points.  `base_points` is the point tensor or, to amortise the grid
    build over the views of a data set, an `ops.KnnGrid` made from it; NERFAIL_B200_KNN=brute selects the plain scan."""
    if isinstance(base_points, ops.KnnGrid):
        return base_points.query_dist_idx(query_hw3)
    if os.environ.get("NERFAIL_B200_KNN", "grid").lower() == "brute":
        return ops.knn8_dist_idx(query_hw3, base_points)
    return ops.KnnGrid(base_points).query_dist_idx(query_hw3)
