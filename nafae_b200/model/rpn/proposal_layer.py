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


def proposal_front(rpn_cls_prob, rpn_bbox_pred, im_info, anchors, feat_stride, pre_nms_topN, return_order=False):
    """Everything `_ProposalLayer.forward` does before its per-frame loop (proposal_layer.py:66-125) in
    two launches: anchors + bbox_transform_inv + clip_boxes + per-frame descending score sort.
    rpn_cls_prob (B, 2A, H, W), rpn_bbox_pred (B, 4A, H, W): CUDA f32; im_info (B, 3); anchors (A, 4).
    Returns (proposals (B, m, 4), scores (B, m)[, order (B, m) int32]) sorted by score, ready for
    `proposal_tail`."""
    cls = _C.f32c(rpn_cls_prob, "rpn_cls_prob")
    dl = _C.f32c(rpn_bbox_pred, "rpn_bbox_pred")
    dev = cls.device
    B, A2, H, W = cls.shape
    A = A2 // 2
    if dl.shape != (B, 4 * A, H, W):
        raise ValueError("rpn_bbox_pred must be (B, 4A, H, W) matching rpn_cls_prob (B, 2A, H, W)")
    info = torch.as_tensor(im_info, dtype=torch.float32).to(dev).contiguous().view(B, -1)
    anc = torch.as_tensor(anchors, dtype=torch.float32).to(dev).contiguous()
    n = H * W * A
    pre = int(pre_nms_topN)
    m = pre if (0 < pre < B * n and pre < n) else n
    props = torch.empty((B, m, 4), dtype=torch.float32, device=dev)
    scrs = torch.empty((B, m), dtype=torch.float32, device=dev)
    order = torch.empty((B, m), dtype=torch.int32, device=dev) if return_order else None
    nbytes = int(_C.lib.nafae_proposal_front_workspace_bytes(B, A, H, W))
    ws = torch.empty((nbytes // 4,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        st = _C.lib.nafae_proposal_front(_C.ptr(cls), _C.ptr(dl), _C.ptr(info), _C.ptr(anc), B, A, H, W,
                                         float(feat_stride), pre, _C.ptr(props), _C.ptr(scrs), _C.ptr(order),
                                         _C.ptr(ws), nbytes, _C.stream(dev))
    _C.check(st, "nafae_proposal_front")
    return (props, scrs, order) if return_order else (props, scrs)


class _ProposalLayer(torch.nn.Module):
    """Mirror of the reference `_ProposalLayer` (lib/model/rpn/proposal_layer.py:26-165): anchors +
    deltas -> boxes -> clip -> per-frame score sort -> NMS -> top-N -> zero-padded `(B, N, 5)` rois.

    Same constructor (`feat_stride, scales, ratios`) and `forward(input)` with
    `input = (rpn_cls_prob, rpn_bbox_pred, im_info, cfg_key)`; the TEST/TRAIN thresholds come from a
    `cfg`-like object (`cfg[cfg_key].RPN_PRE_NMS_TOP_N` ...) passed at construction instead of the
    reference's module-global.  The whole forward is three kernel launches: `proposal_front`
    (decode + clip, then the per-frame sort) and `proposal_tail` (NMS + top-N + padding for all
    frames); no per-frame Python loop, no eager tensor arithmetic, no host synchronisation.
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
        rpn_cls_prob, bbox_deltas, im_info, cfg_key = input[0], input[1], input[2], input[3]
        c = getattr(self._cfg, cfg_key) if not isinstance(self._cfg, dict) else self._cfg[cfg_key]
        pre_nms_topN, post_nms_topN, nms_thresh = c.RPN_PRE_NMS_TOP_N, c.RPN_POST_NMS_TOP_N, c.RPN_NMS_THRESH
        if self._anchors.device != rpn_cls_prob.device:
            self._anchors = self._anchors.to(rpn_cls_prob.device)
        props, scrs = proposal_front(rpn_cls_prob, bbox_deltas, im_info, self._anchors, self._feat_stride,
                                     pre_nms_topN)
        output, self.roi_scores = proposal_tail(props, scrs, 0, post_nms_topN, nms_thresh)
        return output

    def get_roi_score(self):
        return self.roi_scores
