"""Mirror of reference lib/model/nms/nms_wrapper.py:11-18."""
from .nms_gpu import nms_gpu, nms_batched  # noqa: F401


def nms(dets, thresh, force_cpu=False):
    """Greedy NMS of score-sorted boxes; returns int32 CUDA (k, 1) keep indices.

    ``force_cpu`` is accepted and ignored exactly like the reference (nms_wrapper.py:11-18:
    it always calls nms_gpu).  Empty input returns ``[]`` (nms_wrapper.py:13-14).
    """
    if dets.shape[0] == 0:
        return []
    return nms_gpu(dets, thresh)
