"""Mirror of reference lib/model/nms/nms_gpu.py:7-12 on top of the sm_100a batched NMS."""
import torch

from ... import _C


def nms_gpu(dets, thresh):
    """dets: CUDA f32 (n, >=4), rows sorted by score descending -> keep int32 CUDA (k, 1).

    Same contract as the reference (nms_gpu.py:7-12): ``keep[:num_out[0]]``.  Like there, the
    slice needs ``num_out`` on the host, i.e. one device->host sync per call; use
    ``nms_batched`` / ``proposal_tail`` on the hot path to avoid it.
    """
    keep, num_out = nms_batched(dets.unsqueeze(0), thresh)
    return keep[0, : int(num_out[0])].view(-1, 1)


def nms_batched(dets, thresh):
    """dets: CUDA f32 (F, n, dim>=4) -> (keep int32 (F, n), num_out int32 (F,)); no host sync.

    Row f of ``keep`` holds ``num_out[f]`` valid indices (ascending = score order); the rest of
    the row is unspecified.
    """
    dets = _C.f32c(dets, "dets")
    F, n, dim = dets.shape
    keep = torch.empty((F, n), dtype=torch.int32, device=dets.device)
    num_out = torch.zeros((F,), dtype=torch.int32, device=dets.device)
    if F == 0 or n == 0:
        return keep, num_out
    ws_bytes = int(_C.lib.nafae_nms_workspace_bytes(F, n))
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dets.device)
    with torch.cuda.device(dets.device):
        st = _C.lib.nafae_nms_batched(_C.ptr(keep), _C.ptr(num_out), _C.ptr(dets), F, n, dim,
                                      float(thresh), _C.ptr(ws), ws_bytes, _C.stream(dets.device))
    _C.check(st, "nafae_nms_batched")
    return keep, num_out
