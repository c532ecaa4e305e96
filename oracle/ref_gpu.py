"""ctypes wrapper of oracle/_ref/libnafae_ref.so: the reference's own UNMODIFIED CUDA sources
(lib/model/{nms,roi_align,roi_pooling}/src/*.cu) compiled for sm_100a by oracle/Makefile.
TEST INFRASTRUCTURE ONLY -- GPU box only.  Takes / returns torch CUDA tensors."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libnafae_ref.so")
_lib = None
_vp, _i, _f = ctypes.c_void_p, ctypes.c_int, ctypes.c_float


def available():
    return os.path.exists(SO)


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(SO)
        _lib.nms_cuda_compute.restype = None
        _lib.nms_cuda_compute.argtypes = [_vp, _vp, _vp, _i, _i, _f]
        _lib.ROIAlignForwardLaucher.argtypes = [_vp, _f, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]
        _lib.ROIAlignBackwardLaucher.argtypes = [_vp, _f, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]
        _lib.ROIPoolForwardLaucher.argtypes = [_vp, _f, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]
        _lib.ROIPoolBackwardLaucher.argtypes = [_vp, _f, _i, _i, _i, _i, _i, _i, _i, _vp, _vp,
                                                _vp, _vp]
    return _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _s():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def nms(dets, thresh):
    """nms_gpu.py:7-12 driven through nms_cuda_compute (default stream, blocking)."""
    dets = dets.contiguous()
    n = dets.shape[0]
    keep = torch.zeros((n,), dtype=torch.int32, device=dets.device)
    num = torch.zeros((1,), dtype=torch.int32, device=dets.device)
    torch.cuda.synchronize()
    lib().nms_cuda_compute(_p(keep), _p(num), _p(dets), n, dets.shape[1], float(thresh))
    torch.cuda.synchronize()
    return keep[: int(num[0])]


def roi_align_forward(features, rois, ah, aw, scale):
    B, C, H, W = features.shape
    R = rois.shape[0]
    out = torch.zeros((R, C, ah, aw), dtype=torch.float32, device=features.device)
    lib().ROIAlignForwardLaucher(_p(features), float(scale), R, H, W, C, ah, aw, _p(rois), _p(out),
                                 _s())
    return out


def roi_align_backward(top_diff, rois, feature_size, scale):
    B, C, H, W = feature_size
    R, _, ah, aw = top_diff.shape
    bd = torch.zeros((B, C, H, W), dtype=torch.float32, device=top_diff.device)
    lib().ROIAlignBackwardLaucher(_p(top_diff.contiguous()), float(scale), B, R, H, W, C, ah, aw,
                                  _p(rois), _p(bd), _s())
    return bd


def roi_pool_forward(features, rois, ph, pw, scale):
    B, C, H, W = features.shape
    R = rois.shape[0]
    out = torch.zeros((R, C, ph, pw), dtype=torch.float32, device=features.device)
    am = torch.zeros((R, C, ph, pw), dtype=torch.int32, device=features.device)
    lib().ROIPoolForwardLaucher(_p(features), float(scale), R, H, W, C, ph, pw, _p(rois), _p(out),
                                _p(am), _s())
    return out, am


def roi_pool_backward(top_diff, argmax, rois, feature_size, scale):
    B, C, H, W = feature_size
    R, _, ph, pw = top_diff.shape
    bd = torch.zeros((B, C, H, W), dtype=torch.float32, device=top_diff.device)
    lib().ROIPoolBackwardLaucher(_p(top_diff.contiguous()), float(scale), B, R, H, W, C, ph, pw,
                                 _p(rois), _p(bd), _p(argmax), _s())
    return bd
