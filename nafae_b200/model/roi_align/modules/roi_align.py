"""Mirror of reference lib/model/roi_align/modules/roi_align.py:6-42 (same class names, same
constructor arguments, same ``forward(features, rois)``).

``exact=True`` (extension, default False) selects the reference-order mixed fp32/fp64 arithmetic:
bit-identical to the reference kernel + ATen pool, slower.  The default fp32 path is within 1e-4
relative (measured ~1e-6) and is the bandwidth-bound kernel.
"""
from torch.nn.modules.module import Module

from .... import _C
from ..functions.roi_align import RoIAlignFunction


class _Base(Module):
    _pool_mode = _C.POOL_NONE

    def __init__(self, aligned_height, aligned_width, spatial_scale, exact=False):
        super(_Base, self).__init__()
        self.aligned_width = int(aligned_width)
        self.aligned_height = int(aligned_height)
        self.spatial_scale = float(spatial_scale)
        self.exact = bool(exact)

    def forward(self, features, rois):
        return RoIAlignFunction(self.aligned_height, self.aligned_width, self.spatial_scale,
                                self._pool_mode, self.exact)(features, rois)


class RoIAlign(_Base):
    """(R, C, aligned_height, aligned_width) corner-grid samples (ref :6-16)."""
    _pool_mode = _C.POOL_NONE


class RoIAlignAvg(_Base):
    """Samples an (h+1) x (w+1) grid and averages 2x2 windows, stride 1 (ref :18-29) -- fused."""
    _pool_mode = _C.POOL_AVG


class RoIAlignMax(_Base):
    """Samples an (h+1) x (w+1) grid and takes 2x2 window maxima, stride 1 (ref :31-42) -- fused."""
    _pool_mode = _C.POOL_MAX
