"""Fused proposal front end (csrc/proposal_front.cu) against fixtures produced by the reference's OWN
generate_anchors / bbox_transform_inv / clip_boxes (tests/golden/make_front_golden.py cuts them out of
the reference source and executes them unchanged).  Order of the sorted proposals: exact, ties
included (stable descending sort); sorted scores: exact; decoded boxes: the reference arithmetic ran on
the CPU there, whose exp() may differ from CUDA's expf by one ulp -> 2e-6 relative + 2e-4 pixels;
against the same torch ops run on THIS GPU the boxes are bit-identical."""
import glob
import importlib.util
import os

import numpy as np
import pytest
import torch

from _golden import GOLDEN

gpu = pytest.mark.gpu
FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "front_*.npz")))


def _gen():
    spec = importlib.util.spec_from_file_location("make_front_golden", os.path.join(GOLDEN, "make_front_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_anchor_table_of_the_fixtures_is_this_packages():
    from nafae_b200.model.rpn.generate_anchors import generate_anchors
    ours = generate_anchors(scales=np.array([4, 8, 16, 32]), ratios=np.array([0.5, 1, 2])).astype(np.float32)
    for path in FIXTURES:
        np.testing.assert_array_equal(np.load(path)["anchors"], ours)


@gpu
@pytest.mark.parametrize("path", FIXTURES, ids=lambda p: os.path.basename(p)[6:-4])
def test_front_end_matches_reference_functions(path):
    from nafae_b200.model.rpn.proposal_layer import proposal_front
    from nafae_b200.model.rpn.bbox_transform import bbox_transform_inv, clip_boxes
    z = np.load(path)
    B, H, W = int(z["B"]), int(z["H"]), int(z["W"])
    cls_prob, deltas, im_info = _gen().make_case(int(z["seed"]), B, H, W, int(z["img_h"]), int(z["img_w"]),
                                                 bool(z["ties"]))
    dev = torch.device("cuda:0")
    anchors = torch.from_numpy(z["anchors"])
    n = H * W * 12
    props, scores, order = proposal_front(cls_prob.to(dev), deltas.to(dev), im_info, anchors, 16, n,
                                          return_order=True)  # pre_nms_topN = everything: the whole order is checked
    torch.cuda.synchronize()
    assert props.shape == (B, n, 4)
    k = z["order"].shape[1]
    np.testing.assert_array_equal(order.cpu().numpy()[:, :k], z["order"])                 # ties included
    np.testing.assert_array_equal(scores.cpu().numpy()[:, :k], z["scores_sorted"])
    # decoded + clipped boxes (fixture holds every `stride`-th anchor)
    stride = int(z["stride"])
    unsorted = torch.empty_like(props)
    unsorted.scatter_(1, order.long().unsqueeze(2).expand(-1, -1, 4), props)
    np.testing.assert_allclose(unsorted.cpu().numpy()[:, ::stride], z["proposals"], rtol=2e-6, atol=2e-4)
    # ... and bit-identical to the same torch arithmetic on this GPU
    A = 12
    sx = torch.arange(0, W, device=dev, dtype=torch.float32) * 16
    sy = torch.arange(0, H, device=dev, dtype=torch.float32) * 16
    yy, xx = torch.meshgrid(sy, sx, indexing="ij")
    shifts = torch.stack((xx.reshape(-1), yy.reshape(-1), xx.reshape(-1), yy.reshape(-1)), 1)
    anc = (anchors.to(dev).view(1, A, 4) + shifts.view(-1, 1, 4)).view(1, -1, 4).expand(B, -1, 4)
    dl = deltas.to(dev).permute(0, 2, 3, 1).contiguous().view(B, -1, 4)
    want = clip_boxes(bbox_transform_inv(anc, dl, B), im_info.to(dev), B)
    assert torch.equal(unsorted, want)
    # every frame is a permutation
    assert (torch.sort(order.long(), 1)[0] == torch.arange(n, device=dev)).all()


@gpu
def test_pre_nms_topn_and_argument_checks():
    from nafae_b200 import _C
    from nafae_b200.model.rpn.proposal_layer import proposal_front
    g = _gen()
    cls_prob, deltas, im_info = g.make_case(9, 2, 14, 14, 224, 224, False)
    dev = torch.device("cuda:0")
    from nafae_b200.model.rpn.generate_anchors import generate_anchors
    anchors = torch.from_numpy(generate_anchors(scales=np.array([4, 8, 16, 32]), ratios=np.array([0.5, 1, 2]))).float()
    full = proposal_front(cls_prob.to(dev), deltas.to(dev), im_info, anchors, 16, 6000)
    top = proposal_front(cls_prob.to(dev), deltas.to(dev), im_info, anchors, 16, 300)
    assert top[0].shape == (2, 300, 4) and torch.equal(top[0], full[0][:, :300]) and torch.equal(top[1], full[1][:, :300])
    assert _C.lib.nafae_proposal_front_workspace_bytes(2, 12, 14, 14) >= 2 * 2352 * 20
    assert _C.lib.nafae_proposal_front(None, None, None, None, 2, 12, 14, 14, 16.0, 6000, None, None, None, None, 0,
                                       None) == 0
    big = torch.zeros((1, 24, 60, 60), device=dev)  # 43 200 anchors: beyond the shared-memory sort
    with pytest.raises(_C.NafaeError, match="exceed"):
        proposal_front(big, torch.zeros((1, 48, 60, 60), device=dev), torch.tensor([[960., 960., 1.]]), anchors, 16, 6000)
