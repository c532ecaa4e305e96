"""Autograd binding of the sm_100a RoIAlign kernels.

Mirror of reference lib/model/roi_align/functions/roi_align.py:7-47 (``RoIAlignFunction``), which is
a legacy instance-style ``Function`` (``RoIAlignFunction(h, w, s)(features, rois)``) that torch >= 1.5
rejects.  The same call shape is kept on top of a static ``torch.autograd.Function``; the pooled
variants (RoIAlignAvg / RoIAlignMax) are ONE fused kernel instead of kernel + avg/max_pool2d.
"""
import torch

from .... import _C


class _RoIAlign(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, rois, out_h, out_w, spatial_scale, pool_mode, exact):
        features = _C.f32c(features, "features")  # CPU input -> NotImplementedError (ref :28-29)
        rois = _C.f32c(rois, "rois")
        if features.dim() != 4:
            raise ValueError("features must be (B, C, H, W)")
        if rois.dim() != 2 or rois.size(1) != 5:
            # the reference's C glue returns 0 here and leaves the zero-filled output
            # (roi_align_cuda.c:19-22); failing loudly is the only sane drop-in behaviour
            raise ValueError("rois must be (R, 5) [batch_idx, x1, y1, x2, y2]")
        B, C, H, W = features.shape
        R = rois.size(0)
        out = torch.empty((R, C, out_h, out_w), dtype=torch.float32, device=features.device)
        flags = _C.FLAG_EXACT if exact else 0
        with torch.cuda.device(features.device):
            st = _C.lib.nafae_roi_align_forward(
                _C.ptr(features), float(spatial_scale), B, R, H, W, C, int(out_h), int(out_w),
                int(pool_mode), _C.ptr(rois), _C.ptr(out), flags, None, 0,
                _C.stream(features.device))
        _C.check(st, "nafae_roi_align_forward")
        ctx.cfg = (int(out_h), int(out_w), float(spatial_scale), int(pool_mode), flags)
        ctx.feature_size = (B, C, H, W)
        if pool_mode == _C.POOL_MAX:
            ctx.save_for_backward(rois, features)
        else:
            ctx.save_for_backward(rois)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_output):
        out_h, out_w, scale, pool_mode, flags = ctx.cfg
        B, C, H, W = ctx.feature_size
        saved = ctx.saved_tensors
        rois = saved[0]
        features = saved[1] if pool_mode == _C.POOL_MAX else None
        grad_output = _C.f32c(grad_output, "grad_output")
        # the reference zero-fills and accumulates with atomics (functions/roi_align.py:38-39); here the
        # call overwrites grad_input (NAFAE_FLAG_OVERWRITE): no memset, and RoIAlignAvg 7x7 takes the
        # atomic-free cell-gather kernel
        grad_input = torch.empty((B, C, H, W), dtype=torch.float32, device=grad_output.device)
        with torch.cuda.device(grad_output.device):
            st = _C.lib.nafae_roi_align_backward(
                _C.ptr(grad_output), _C.ptr(features), scale, B, rois.size(0), H, W, C, out_h,
                out_w, pool_mode, _C.ptr(rois), _C.ptr(grad_input), flags | _C.FLAG_OVERWRITE,
                _C.stream(grad_output.device))
        _C.check(st, "nafae_roi_align_backward")
        return grad_input, None, None, None, None, None, None  # (grad, None) as ref :47


class RoIAlignFunction(object):
    """``RoIAlignFunction(aligned_height, aligned_width, spatial_scale)(features, rois)``."""

    def __init__(self, aligned_height, aligned_width, spatial_scale, pool_mode=_C.POOL_NONE,
                 exact=False):
        self.aligned_width = int(aligned_width)
        self.aligned_height = int(aligned_height)
        self.spatial_scale = float(spatial_scale)
        self.pool_mode = pool_mode
        self.exact = bool(exact)

    def __call__(self, features, rois):
        return _RoIAlign.apply(features, rois, self.aligned_height, self.aligned_width,
                               self.spatial_scale, self.pool_mode, self.exact)
