"""ctypes front end of the C restatement (oracle/nafae_oracle.c).  TEST INFRASTRUCTURE ONLY.

All functions take / return numpy arrays (float32 / int32, C-contiguous) and mirror the
reference operator surface they check:

* ``nms``            lib/model/nms/nms_wrapper.py:11-18 (+ nms_cuda_kernel.cu:31-144)
* ``proposal_tail``  lib/model/rpn/proposal_layer.py:127-163
* ``roi_align*``     lib/model/roi_align/{modules,functions}/roi_align.py, src/roi_align_kernel.cu
* ``roi_pool*``      lib/model/roi_pooling/functions/roi_pool.py, src/roi_pooling_kernel.cu
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libnafae_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int)
_u64p = ctypes.POINTER(ctypes.c_uint64)


def build(force=False):
    """Compile nafae_oracle.c (and, when /root/reference exists, oracle/_ref)."""
    src = os.path.join(_HERE, "nafae_oracle.c")
    stale = (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "libnafae_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_num_threads.restype = ctypes.c_int
        _lib.oracle_iou_pair.restype = ctypes.c_float
    return _lib


def num_threads():
    return int(lib().oracle_num_threads())


def set_num_threads(n):
    lib().oracle_set_num_threads(ctypes.c_int(int(n)))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t):
    return a.ctypes.data_as(t)


def iou(a, b):
    a, b = _f32(a), _f32(b)
    return float(lib().oracle_iou_pair(_p(a, _f32p), _p(b, _f32p)))


def nms(dets, thresh, return_mask=False):
    """dets (n, >=4) sorted by score desc -> keep indices int32 (k,), ascending."""
    dets = _f32(dets)
    n = dets.shape[0]
    if n == 0:
        return (np.zeros((0,), np.int32), None) if return_mask else np.zeros((0,), np.int32)
    dim = dets.shape[1]
    keep = np.zeros((n,), np.int32)
    num = np.zeros((1,), np.int32)
    mask = np.zeros((n, (n + 63) // 64), np.uint64) if return_mask else None
    lib().oracle_nms(_p(dets, _f32p), ctypes.c_int(n), ctypes.c_int(dim), ctypes.c_float(thresh),
                     _p(keep, _i32p), _p(num, _i32p),
                     _p(mask, _u64p) if return_mask else None)
    keep = keep[: int(num[0])].copy()
    return (keep, mask) if return_mask else keep


def proposal_tail(proposals, scores, pre_nms_topn, post_nms_topn, thresh):
    """proposals (F,n,4), scores (F,n) sorted desc per frame -> rois (F,post,5), roi_scores
    (F,post), num_kept (F,)."""
    proposals, scores = _f32(proposals), _f32(scores)
    F, n = scores.shape
    rois = np.zeros((F, post_nms_topn, 5), np.float32)
    rsc = np.zeros((F, post_nms_topn), np.float32)
    nk = np.zeros((F,), np.int32)
    lib().oracle_proposal_tail(_p(proposals, _f32p), _p(scores, _f32p), ctypes.c_int(F),
                               ctypes.c_int(n), ctypes.c_int(pre_nms_topn),
                               ctypes.c_int(post_nms_topn), ctypes.c_float(thresh),
                               _p(rois, _f32p), _p(rsc, _f32p), _p(nk, _i32p))
    return rois, rsc, nk


def roi_align_forward(features, rois, ah, aw, scale):
    """RoIAlignFunction.forward: (B,C,H,W),(R,5) -> (R,C,ah,aw)."""
    features, rois = _f32(features), _f32(rois)
    B, C, H, W = features.shape
    R = rois.shape[0]
    top = np.zeros((R, C, ah, aw), np.float32)
    lib().oracle_roi_align_forward(_p(features, _f32p), ctypes.c_float(scale), ctypes.c_int(R),
                                   ctypes.c_int(H), ctypes.c_int(W), ctypes.c_int(C),
                                   ctypes.c_int(ah), ctypes.c_int(aw), _p(rois, _f32p),
                                   _p(top, _f32p))
    return top


def roi_align_backward(top_diff, rois, feature_size, scale):
    """RoIAlignFunction.backward: (R,C,ah,aw) -> (B,C,H,W)."""
    top_diff, rois = _f32(top_diff), _f32(rois)
    B, C, H, W = feature_size
    R, _, ah, aw = top_diff.shape
    bd = np.zeros((B, C, H, W), np.float32)
    lib().oracle_roi_align_backward(_p(top_diff, _f32p), ctypes.c_float(scale), ctypes.c_int(B),
                                    ctypes.c_int(R), ctypes.c_int(H), ctypes.c_int(W),
                                    ctypes.c_int(C), ctypes.c_int(ah), ctypes.c_int(aw),
                                    _p(rois, _f32p), _p(bd, _f32p))
    return bd


def pool2x2_forward(x, is_max):
    x = _f32(x)
    ih, iw = x.shape[-2:]
    planes = int(np.prod(x.shape[:-2]))
    y = np.zeros(x.shape[:-2] + (ih - 1, iw - 1), np.float32)
    lib().oracle_pool2x2_forward(_p(x, _f32p), ctypes.c_int(planes), ctypes.c_int(ih),
                                 ctypes.c_int(iw), ctypes.c_int(int(is_max)), _p(y, _f32p))
    return y


def pool2x2_backward(x, gy, is_max):
    x, gy = _f32(x), _f32(gy)
    ih, iw = x.shape[-2:]
    planes = int(np.prod(x.shape[:-2]))
    gx = np.zeros_like(x)
    lib().oracle_pool2x2_backward(_p(x, _f32p), _p(gy, _f32p), ctypes.c_int(planes),
                                  ctypes.c_int(ih), ctypes.c_int(iw), ctypes.c_int(int(is_max)),
                                  _p(gx, _f32p))
    return gx


def roi_align_avg_forward(features, rois, ah, aw, scale):
    """RoIAlignAvg.forward (modules/roi_align.py:26-29): sample (ah+1)x(aw+1), avg 2x2."""
    return pool2x2_forward(roi_align_forward(features, rois, ah + 1, aw + 1, scale), False)


def roi_align_max_forward(features, rois, ah, aw, scale):
    """RoIAlignMax.forward (modules/roi_align.py:39-42)."""
    return pool2x2_forward(roi_align_forward(features, rois, ah + 1, aw + 1, scale), True)


def roi_align_avg_backward(grad_out, features, rois, scale):
    x = roi_align_forward(features, rois, grad_out.shape[2] + 1, grad_out.shape[3] + 1, scale)
    gx = pool2x2_backward(x, grad_out, False)
    return roi_align_backward(gx, rois, features.shape, scale)


def roi_align_max_backward(grad_out, features, rois, scale):
    x = roi_align_forward(features, rois, grad_out.shape[2] + 1, grad_out.shape[3] + 1, scale)
    gx = pool2x2_backward(x, grad_out, True)
    return roi_align_backward(gx, rois, features.shape, scale)


def roi_pool_forward(features, rois, ph, pw, scale):
    """RoIPoolFunction.forward -> (output (R,C,ph,pw) f32, argmax int32)."""
    features, rois = _f32(features), _f32(rois)
    B, C, H, W = features.shape
    R = rois.shape[0]
    top = np.zeros((R, C, ph, pw), np.float32)
    am = np.zeros((R, C, ph, pw), np.int32)
    lib().oracle_roi_pool_forward(_p(features, _f32p), ctypes.c_float(scale), ctypes.c_int(R),
                                  ctypes.c_int(H), ctypes.c_int(W), ctypes.c_int(C),
                                  ctypes.c_int(ph), ctypes.c_int(pw), _p(rois, _f32p),
                                  _p(top, _f32p), _p(am, _i32p))
    return top, am


def roi_pool_backward(top_diff, argmax, rois, feature_size, scale):
    top_diff, rois = _f32(top_diff), _f32(rois)
    argmax = np.ascontiguousarray(argmax, dtype=np.int32)
    B, C, H, W = feature_size
    R, _, ph, pw = top_diff.shape
    bd = np.zeros((B, C, H, W), np.float32)
    lib().oracle_roi_pool_backward(_p(top_diff, _f32p), _p(argmax, _i32p), ctypes.c_float(scale),
                                   ctypes.c_int(B), ctypes.c_int(R), ctypes.c_int(H),
                                   ctypes.c_int(W), ctypes.c_int(C), ctypes.c_int(ph),
                                   ctypes.c_int(pw), _p(rois, _f32p), _p(bd, _f32p))
    return bd
