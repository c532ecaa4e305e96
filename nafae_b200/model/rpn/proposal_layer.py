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
