"""Host-side mirror of the reference's dense-energy (CRF regularisation) loss for the SCD-AAAI2023 stack, SURVEY.md 8(f) rank 4:

  * `bilateralfilter_batch(images, ins, outs, N, K, H, W, sigma_rgb, sigma_xy)` — the SWIG module function of
    SCD-AAAI2023/wrapper/bilateralfilter/bilateralfilter.py (C++: bilateralfilter.hpp:12), same arguments, flat fp32 numpy arrays,
    `outs` written in place; runs `rss_bilateralfilter_batch_host` (H2D, device filter, D2H);
  * `bilateral_filter(images, ins, sigma_rgb, sigma_xy)` — the same filter on DEVICE tensors, stream-ordered, no host round trip;
  * `DenseEnergyLossFunction` / `DenseEnergyLoss(weight, sigma_rgb, sigma_xy, scale_factor)` — SCD-AAAI2023/utils/losses.py:52-120
    with the same constructor, `forward(images, segmentations, ROIs, seg_label)` and gradient, but the tensors never leave the
    GPU (the reference copies images and segmentations to the host, filters there with OpenMP, and copies the result back
    every training step: losses.py:66-76,83-85).

No CPU fallback: every call goes through librss_b200.so and raises without it / off sm_100.
"""
import ctypes

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib

_FP = ctypes.POINTER(ctypes.c_float)


def bilateralfilter_batch(images, ins, outs, N, K, H, W, sigma_rgb, sigma_xy):
    """drop-in for `from bilateralfilter import bilateralfilter_batch` (utils/losses.py:8,70): host numpy arrays, in-place outs"""
    for name, a in (("images", images), ("ins", ins), ("outs", outs)):
        if not (isinstance(a, np.ndarray) and a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]):
            raise TypeError("%s must be a C-contiguous float32 numpy array (the SWIG typemap's requirement)" % name)
    _lib.require_device()
    rc = _lib.load().rss_bilateralfilter_batch_host(images.ctypes.data, images.size, ins.ctypes.data, ins.size, outs.ctypes.data, outs.size,
                                                    N, K, H, W, float(sigma_rgb), float(sigma_xy))
    _lib.check(rc, "rss_bilateralfilter_batch_host")


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def bilateral_filter(images, ins, sigma_rgb, sigma_xy, out=None, lattice_points=None):
    """images (N,3,H,W), ins (N,K,H,W): contiguous fp32 CUDA tensors -> filtered (N,K,H,W).  lattice_points: optional int32 CUDA
    tensor [1] receiving the number of lattice points of the batch."""
    _lib.require_device()
    if not (images.is_cuda and ins.is_cuda and images.dtype == torch.float32 and ins.dtype == torch.float32):
        raise TypeError("bilateral_filter wants fp32 CUDA tensors")
    images, ins = images.contiguous(), ins.contiguous()
    N, K, H, W = ins.shape
    if tuple(images.shape) != (N, 3, H, W):
        raise ValueError("images must be (N,3,H,W) matching ins (N,K,H,W)")
    lib = _lib.load()
    nbytes = lib.rss_bilateral_workspace_bytes(N, K, H, W)
    if nbytes == 0:
        raise ValueError("unsupported bilateral filter shape %r" % ((N, K, H, W),))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=ins.device)
    if out is None:
        out = torch.empty_like(ins)
    rc = lib.rss_bilateralfilter_batch(images.data_ptr(), ins.data_ptr(), out.data_ptr(), N, K, H, W, float(sigma_rgb), float(sigma_xy),
                                       ws.data_ptr(), nbytes, lattice_points.data_ptr() if lattice_points is not None else None, _stream())
    _lib.check(rc, "rss_bilateralfilter_batch")
    return out


class DenseEnergyLossFunction(torch.autograd.Function):
    """utils/losses.py:52-91.  forward: gate = clamp(ROI - max_k seg, 0) (1 where unlabeled); AS = gate * filter(images, seg * ROI);
    loss = -<seg * ROI, AS> / N.  backward: d seg = -2 * grad * AS / N * ROI (the reference's approximation: the gate is constant)."""

    @staticmethod
    def forward(ctx, images, segmentations, sigma_rgb, sigma_xy, ROIs, unlabel_region):
        N, K, H, W = segmentations.shape
        seg = segmentations.detach().float().contiguous()
        rois = ROIs.detach().float().contiguous()
        seg_roi = seg * rois.unsqueeze(1)
        AS = bilateral_filter(images.detach().float(), seg_roi, sigma_rgb, sigma_xy)
        loss = torch.zeros(1, dtype=torch.float64, device=seg.device)
        unl = unlabel_region.to(torch.uint8).contiguous()
        rc = _lib.load().rss_dense_energy_gate(seg.data_ptr(), rois.data_ptr(), unl.data_ptr(), seg_roi.data_ptr(), AS.data_ptr(),
                                               loss.data_ptr(), N, K, H, W, _stream())
        _lib.check(rc, "rss_dense_energy_gate")
        ctx.N = N
        ctx.save_for_backward(AS, rois)
        return loss.float()

    @staticmethod
    def backward(ctx, grad_output):
        AS, rois = ctx.saved_tensors
        grad_segmentation = torch.mul(-2 * grad_output * AS / ctx.N, rois.unsqueeze(1))      # losses.py:87-89, same operation order
        return None, grad_segmentation, None, None, None, None


class DenseEnergyLoss(nn.Module):
    """utils/losses.py:94-120: same constructor and forward signature"""

    def __init__(self, weight, sigma_rgb, sigma_xy, scale_factor):
        super().__init__()
        self.weight = weight
        self.sigma_rgb = sigma_rgb
        self.sigma_xy = sigma_xy
        self.scale_factor = scale_factor

    def forward(self, images, segmentations, ROIs, seg_label):
        scaled_images = F.interpolate(images, scale_factor=self.scale_factor)
        scaled_segs = F.interpolate(segmentations, scale_factor=self.scale_factor, mode="bilinear", align_corners=False)
        scaled_ROIs = F.interpolate(ROIs.unsqueeze(1), scale_factor=self.scale_factor).squeeze(1)
        scaled_seg_label = F.interpolate(seg_label, scale_factor=self.scale_factor, mode="nearest")
        unlabel_region = (scaled_seg_label.long() == 255).squeeze(1)
        return self.weight * DenseEnergyLossFunction.apply(scaled_images, scaled_segs, self.sigma_rgb, self.sigma_xy * self.scale_factor,
                                                           scaled_ROIs, unlabel_region)

    def extra_repr(self):
        return "sigma_rgb={}, sigma_xy={}, weight={}, scale_factor={}".format(self.sigma_rgb, self.sigma_xy, self.weight, self.scale_factor)
