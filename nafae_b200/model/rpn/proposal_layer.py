"""Batched tail of the reference _ProposalLayer.forward (lib/model/rpn/proposal_layer.py:127-163).

The reference runs a Python loop over the frames of the batch; every iteration concatenates
boxes and scores, calls ``nms()`` (two cudaMallocs, a 696 KB mask D2H copy, a host sweep and two
more syncs), slices the first ``post_nms_topN`` keeps and writes them into a zero-padded output.
Here the whole loop is one kernel launch with no host synchronisation.
"""
import torch

from ... import _C


def proposal_tail(proposals, scores, pre_nms_topN, post_nms_topN, nms_thresh, return_num=False):
    """proposals (F, n, 4), scores (F, n): CUDA f32, each frame already sorted by score desc
    (proposal_layer.py:125).  Returns (rois (F, post, 5), roi_scores (F, post)[, num_kept (F,)]).

    rois rows are [frame, x1, y1, x2, y2]; rows past the number of kept boxes are
    [frame, 0, 0, 0, 0] with score 0 (proposal_layer.py:127-129,158-163).
    """
    proposals = _C.f32c(proposals, "proposals")
    scores = _C.f32c(scores, "scores")
    F, n = scores.shape
    if proposals.shape != (F, n, 4):
        raise ValueError("proposals must be (F, n, 4) matching scores (F, n)")
    post = int(post_nms_topN)
    rois = torch.empty((F, post, 5), dtype=torch.float32, device=scores.device)
    rsc = torch.empty((F, post), dtype=torch.float32, device=scores.device)
    num = torch.empty((F,), dtype=torch.int32, device=scores.device) if return_num else None
    with torch.cuda.device(scores.device):
        st = _C.lib.nafae_proposal_tail(_C.ptr(proposals), _C.ptr(scores), F, n,
                                        int(pre_nms_topN), post, float(nms_thresh), _C.ptr(rois),
                                        _C.ptr(rsc), _C.ptr(num), _C.stream(scores.device))
    _C.check(st, "nafae_proposal_tail")
    return (rois, rsc, num) if return_num else (rois, rsc)


class _ProposalLayer(torch.nn.Module):
    """Mirror of the reference `_ProposalLayer` (lib/model/rpn/proposal_layer.py:26-165): anchors +
    deltas -> boxes -> clip -> per-frame score sort -> NMS -> top-N -> zero-padded `(B, N, 5)` rois.

    Same constructor (`feat_stride, scales, ratios`) and `forward(input)` with
    `input = (rpn_cls_prob, rpn_bbox_pred, im_info, cfg_key)`; the TEST/TRAIN thresholds come from a
    `cfg`-like object (`cfg[cfg_key].RPN_PRE_NMS_TOP_N` ...) passed at construction instead of the
    reference's module-global.  Everything up to the sort is the reference's own torch arithmetic
    (:80-125); the Python loop over frames (:130-163) is one `proposal_tail` launch.
    """

    def __init__(self, feat_stride, scales, ratios, cfg):
        super(_ProposalLayer, self).__init__()
        import numpy as np
        from .generate_anchors import generate_anchors
        self._feat_stride = feat_stride
        self._anchors = torch.from_numpy(generate_anchors(scales=np.array(scales),
                                                          ratios=np.array(ratios))).float()
        self._num_anchors = self._anchors.size(0)
        self._cfg = cfg
        self.roi_scores = None

    def forward(self, input):
        from .bbox_transform import bbox_transform_inv, clip_boxes
        scores = input[0][:, self._num_anchors:, :, :]  # fg probabilities
        bbox_deltas, im_info, cfg_key = input[1], input[2], input[3]
        c = getattr(self._cfg, cfg_key) if not isinstance(self._cfg, dict) else self._cfg[cfg_key]
        pre_nms_topN, post_nms_topN, nms_thresh = c.RPN_PRE_NMS_TOP_N, c.RPN_POST_NMS_TOP_N, c.RPN_NMS_THRESH
        batch_size = bbox_deltas.size(0)
        H, W = scores.size(2), scores.size(3)
        dev = scores.device
        sx = torch.arange(0, W, device=dev, dtype=torch.float32) * self._feat_stride
        sy = torch.arange(0, H, device=dev, dtype=torch.float32) * self._feat_stride
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        shifts = torch.stack((xx.reshape(-1), yy.reshape(-1), xx.reshape(-1), yy.reshape(-1)), 1)
        A, K = self._num_anchors, shifts.size(0)
        anchors = (self._anchors.to(dev).view(1, A, 4) + shifts.view(K, 1, 4)).view(1, K * A, 4)
        anchors = anchors.expand(batch_size, K * A, 4)
        bbox_deltas = bbox_deltas.permute(0, 2, 3, 1).contiguous().view(batch_size, -1, 4)
        scores = scores.permute(0, 2, 3, 1).contiguous().view(batch_size, -1)
        proposals = clip_boxes(bbox_transform_inv(anchors, bbox_deltas, batch_size), im_info, batch_size)
        _, order = torch.sort(scores, 1, True)  # :125
        if 0 < pre_nms_topN < scores.numel():   # :139-140 (numel of the whole batch, as there)
            order = order[:, :pre_nms_topN]
        props = proposals.gather(1, order.unsqueeze(2).expand(-1, -1, 4)).contiguous()
        scrs = scores.gather(1, order).contiguous()
        output, self.roi_scores = proposal_tail(props, scrs, 0, post_nms_topN, nms_thresh)
        return output

    def get_roi_score(self):
        return self.roi_scores
