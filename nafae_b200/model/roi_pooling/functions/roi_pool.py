"""Autograd binding of the sm_100a RoIPool kernels; mirror of reference
lib/model/roi_pooling/functions/roi_pool.py:6-38 (``RoIPoolFunction``)."""
import torch

from .... import _C


class _RoIPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, rois, pooled_height, pooled_width, spatial_scale):
        # the reference's CPU branch (roi_pool.py:20-23, src/roi_pooling.c) is defective
        #  and is not reproduced: CUDA only
        features = _C.f32c(features, "features")
        rois = _C.f32c(rois, "rois")
        if rois.dim() != 2 or rois.size(1) != 5:
            raise ValueError("rois must be (R, 5) [batch_idx, x1, y1, x2, y2]")
        B, C, H, W = features.shape
        R = rois.size(0)
        out = torch.empty((R, C, pooled_height, pooled_width), dtype=torch.float32,
                          device=features.device)
        argmax = torch.empty((R, C, pooled_height, pooled_width), dtype=torch.int32,
                             device=features.device)
        with torch.cuda.device(features.device):
            st = _C.lib.ROIPoolForwardLaucher(_C.ptr(features), float(spatial_scale), R, H, W, C,
                                              int(pooled_height), int(pooled_width), _C.ptr(rois),
                                              _C.ptr(out), _C.ptr(argmax),
                                              _C.stream(features.device))
        _C.check(st, "ROIPoolForwardLaucher")
        ctx.cfg = (int(pooled_height), int(pooled_width), float(spatial_scale))
        ctx.feature_size = (B, C, H, W)
        ctx.save_for_backward(rois, argmax)
        ctx.mark_non_differentiable(argmax)
        return out, argmax

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_output, _grad_argmax):
        ph, pw, scale = ctx.cfg
        B, C, H, W = ctx.feature_size
        rois, argmax = ctx.saved_tensors
        grad_output = _C.f32c(grad_output, "grad_output")
        grad_input = torch.empty((B, C, H, W), dtype=torch.float32, device=grad_output.device)
        with torch.cuda.device(grad_output.device):
            st = _C.lib.ROIPoolBackwardLaucher(_C.ptr(grad_output), scale, B, rois.size(0), H, W,
                                               C, ph, pw, _C.ptr(rois), _C.ptr(grad_input),
                                               _C.ptr(argmax), _C.stream(grad_output.device))
        _C.check(st, "ROIPoolBackwardLaucher")
        return grad_input, None, None, None, None


class RoIPoolFunction(object):
    """``RoIPoolFunction(pooled_height, pooled_width, spatial_scale)(features, rois)``; the argmax
    of the last call is kept on the instance like the reference's ``ctx.argmax`` (roi_pool.py:18)."""

    def __init__(self, pooled_height, pooled_width, spatial_scale):
        self.pooled_width = int(pooled_width)
        self.pooled_height = int(pooled_height)
        self.spatial_scale = float(spatial_scale)
        self.argmax = None

    def __call__(self, features, rois):
        out, self.argmax = _RoIPool.apply(features, rois, self.pooled_height, self.pooled_width,
                                          self.spatial_scale)
        return out
