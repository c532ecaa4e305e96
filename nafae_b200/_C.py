"""ctypes binding of libnafae_b200.so (the C ABI declared in include/nafae_b200.h).

This is the only place the shared library is loaded.  It fails loudly when the library is
missing -- there is deliberately no CPU or PyTorch fallback.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NAFAE_B200_LIB", os.path.join(_HERE, "libnafae_b200.so"))  # override: debug builds

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "nafae_b200: %s not found. Build it with `python -c 'import __graft_entry__ as g; "
        "g.build()'` or `make -C nafae_b200/csrc`. There is no CPU fallback." % LIB_PATH)

lib = ctypes.CDLL(LIB_PATH)

c_int, c_float, c_size_t, c_void_p, c_uint = (ctypes.c_int, ctypes.c_float, ctypes.c_size_t,
                                              ctypes.c_void_p, ctypes.c_uint)

# name -> (restype, argtypes); kept in the order of include/nafae_b200.h
SIGNATURES = {
    "nafae_abi_version": (c_int, []),
    "nafae_last_error": (ctypes.c_char_p, []),
    "nafae_set_reserved_sms": (c_int, [c_int]),
    "nafae_gate_wait": (c_int, [c_void_p, c_int, c_void_p]),
    "nafae_gate_sync": (c_int, [c_void_p, c_void_p]),
    "nafae_debug_occupy_sms": (c_int, [c_int, c_int, ctypes.c_ulonglong, c_void_p]),
    "nms_cuda_compute": (None, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float]),
    "nafae_nms_workspace_bytes": (c_size_t, [c_int, c_int]),
    "nafae_nms_batched": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float,
                                  c_void_p, c_size_t, c_void_p]),
    "nafae_proposal_front_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "nafae_proposal_front": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float,
                                     c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "nafae_proposal_tail": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float,
                                    c_void_p, c_void_p, c_void_p, c_void_p]),
    "ROIAlignForwardLaucher": (c_int, [c_void_p, c_float, c_int, c_int, c_int, c_int, c_int, c_int,
                                       c_void_p, c_void_p, c_void_p]),
    "ROIAlignBackwardLaucher": (c_int, [c_void_p, c_float, c_int, c_int, c_int, c_int, c_int,
                                        c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "nafae_roi_align_workspace_bytes": (c_size_t, [c_int, c_int]),
    "nafae_roi_align_persistent_ctas": (c_int, [c_int]),
    "nafae_roi_align_forward": (c_int, [c_void_p, c_float, c_int, c_int, c_int, c_int, c_int,
                                        c_int, c_int, c_int, c_void_p, c_void_p, c_uint, c_void_p,
                                        c_size_t, c_void_p]),
    "nafae_roi_align_backward": (c_int, [c_void_p, c_void_p, c_float, c_int, c_int, c_int, c_int,
                                         c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_uint,
                                         c_void_p]),
    "ROIPoolForwardLaucher": (c_int, [c_void_p, c_float, c_int, c_int, c_int, c_int, c_int, c_int,
                                      c_void_p, c_void_p, c_void_p, c_void_p]),
    "ROIPoolBackwardLaucher": (c_int, [c_void_p, c_float, c_int, c_int, c_int, c_int, c_int,
                                       c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "nafae_ground_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "nafae_ground_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                     c_int, c_float, c_float, c_int, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_size_t, c_void_p]),
    "nafae_ground_forward_batched": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                             c_int, c_int, c_float, c_float, c_int, c_void_p, c_void_p,
                                             c_void_p, c_void_p, c_size_t, c_void_p]),
    "nafae_ground_forward_tc": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                        c_int, c_float, c_float, c_int, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_size_t, c_void_p]),
    "nafae_ground_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                      c_int, c_int, c_float, c_float, c_int, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "nafae_ground_postprocess": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                         c_void_p, c_void_p]),
    "nafae_eval_record": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                  ctypes.c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_float, c_int, c_void_p, c_void_p, c_void_p]),
    "nafae_gemm_bf16_tn": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_uint, c_void_p]),
    "nafae_clip_adam_workspace_bytes": (c_size_t, []),
    "nafae_clip_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_float, c_float,
                                     c_float, c_float, c_float, c_float, c_void_p, c_size_t, c_void_p]),
    "nafae_ar_buffer_bytes": (c_size_t, [c_size_t, c_int]),
    "nafae_ar_data_offset": (c_size_t, []),
    "nafae_ar_alloc": (c_int, [c_size_t, ctypes.POINTER(c_void_p), c_void_p]),
    "nafae_ar_open": (c_int, [c_void_p, ctypes.POINTER(c_void_p)]),
    "nafae_ar_close": (c_int, [c_void_p]),
    "nafae_ar_free": (c_int, [c_void_p]),
    "nafae_allreduce_avg": (c_int, [c_void_p, c_int, c_int, c_size_t, c_int, c_int, c_uint, c_void_p]),
    "nafae_mc_supported": (c_int, []),
    "nafae_mc_buffer_bytes": (c_size_t, [c_size_t, c_int]),
    "nafae_mc_create": (c_int, [c_int, c_size_t, ctypes.POINTER(c_void_p), ctypes.POINTER(c_int)]),
    "nafae_mc_import": (c_int, [c_int, c_int, c_size_t, ctypes.POINTER(c_void_p)]),
    "nafae_mc_add_device": (c_int, [c_void_p]),
    "nafae_mc_bind": (c_int, [c_void_p, ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p)]),
    "nafae_mc_size": (c_size_t, [c_void_p]),
    "nafae_mc_free": (c_int, [c_void_p]),
    "nafae_allreduce_mc": (c_int, [c_void_p, c_void_p, c_int, c_int, c_size_t, c_int, c_int, c_void_p]),
    "nafae_allreduce_mc_error": (c_int, [c_void_p]),
}

MISSING = []
for _name, (_res, _args) in SIGNATURES.items():
    try:
        _fn = getattr(lib, _name)
    except AttributeError:
        MISSING.append(_name)
        continue
    _fn.restype = _res
    _fn.argtypes = _args

ABI_VERSION = 3  # include/nafae_b200.h NAFAE_B200_ABI_VERSION this host code was written against
if not MISSING and int(lib.nafae_abi_version()) != ABI_VERSION:
    raise ImportError("%s has ABI version %d, this package needs %d: rebuild it (make -C nafae_b200/csrc)"
                      % (LIB_PATH, int(lib.nafae_abi_version()), ABI_VERSION))

POOL_NONE, POOL_AVG, POOL_MAX = 0, 1, 2
FLAG_EXACT = 1
FLAG_NO_GATE = 2
FLAG_OUT_BF16 = 4
FLAG_OVERWRITE = 8
FLAG_DETERMINISTIC = 16
GATE_BYTES = 32
ROI_ALIGN_WS_BYTES = 64


class NafaeError(RuntimeError):
    pass


def last_error():
    msg = lib.nafae_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(status, what):
    """1 = ok; 0 = invalid argument; <0 = -cudaError (include/nafae_b200.h)."""
    if status != 1:
        raise NafaeError("%s failed (status %d): %s" % (what, status, last_error()))


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def stream(device=None):
    """cudaStream_t of torch's current stream on `device`."""
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t, name):
    if not t.is_cuda:
        raise NotImplementedError("%s must be a CUDA tensor: nafae_b200 has no CPU path" % name)


def f32c(t, name):
    require_cuda(t, name)
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32, got %s" % (name, t.dtype))
    return t if t.is_contiguous() else t.contiguous()
