"""Proposal front end, CPU part: the anchor enumeration against the
reference's only golden vector (lib/model/rpn/generate_anchors.py:12-37) and the torch box decoding
against an independent numpy restatement."""
import numpy as np
import torch

from nafae_b200.model.rpn.generate_anchors import generate_anchors
from nafae_b200.model.rpn.bbox_transform import bbox_transform_inv, clip_boxes

GOLDEN_9 = np.array([[-83, -39, 100, 56], [-175, -87, 192, 104], [-359, -183, 376, 200],
                     [-55, -55, 72, 72], [-119, -119, 136, 136], [-247, -247, 264, 264],
                     [-35, -79, 52, 96], [-79, -167, 96, 184], [-167, -343, 184, 360]], np.float64)


# generate_anchors(scales=[4,8,16,32], ratios=[.5,1,2]) of the reference itself, executed in the build
# container (importlib on /root/reference/lib/model/rpn/generate_anchors.py): NAFAE's 12 anchors
REF_12 = np.array([[-38, -16, 53, 31], [-84, -40, 99, 55], [-176, -88, 191, 103], [-360, -184, 375, 199],
                   [-24, -24, 39, 39], [-56, -56, 71, 71], [-120, -120, 135, 135], [-248, -248, 263, 263],
                   [-14, -36, 29, 51], [-36, -80, 51, 95], [-80, -168, 95, 183], [-168, -344, 183, 359]],
                  np.float64)


def test_anchors_match_reference_golden_vector():
    """The table in the reference's comment is the MATLAB (1-based) output it was checked against:
    the Python function returns the same windows in 0-based pixel coordinates, i.e. the table - 1
    (verified by running the reference's own function)."""
    np.testing.assert_array_equal(generate_anchors() + 1, GOLDEN_9)


def test_anchors_match_reference_function_output():
    np.testing.assert_array_equal(generate_anchors(scales=np.array([4, 8, 16, 32]), ratios=np.array([0.5, 1, 2])),
                                  REF_12)


def test_nafae_anchor_configuration():
    """cfgs/vgg16.yml:14-18: scales [4,8,16,32] x ratios [.5,1,2] = 12 anchors; 14x14 map -> 2352 boxes."""
    a = generate_anchors(scales=np.array([4, 8, 16, 32]), ratios=np.array([0.5, 1, 2]))
    assert a.shape == (12, 4) and 14 * 14 * a.shape[0] == 2352
    w, h = a[:, 2] - a[:, 0] + 1, a[:, 3] - a[:, 1] + 1
    np.testing.assert_allclose((w * h).reshape(3, 4) / np.array([4, 8, 16, 32]) ** 2, 256, rtol=0.08)
    np.testing.assert_allclose((a[:, 0] + a[:, 2]) / 2, 7.5)  # all centred on the base window


def test_decode_and_clip_match_numpy():
    rs = np.random.RandomState(0)
    anchors = np.sort(rs.uniform(0, 600, (2, 50, 4)).astype(np.float32).reshape(2, 50, 2, 2), 2)
    anchors = anchors.transpose(0, 1, 3, 2).reshape(2, 50, 4)[..., [0, 2, 1, 3]]
    deltas = (rs.standard_normal((2, 50, 4)) * 0.3).astype(np.float32)
    got = bbox_transform_inv(torch.from_numpy(anchors), torch.from_numpy(deltas), 2)
    w = anchors[..., 2] - anchors[..., 0] + 1
    h = anchors[..., 3] - anchors[..., 1] + 1
    cx, cy = anchors[..., 0] + 0.5 * w, anchors[..., 1] + 0.5 * h
    pcx, pcy = deltas[..., 0] * w + cx, deltas[..., 1] * h + cy
    pw, ph = np.exp(deltas[..., 2]) * w, np.exp(deltas[..., 3]) * h
    want = np.stack([pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph], -1)
    np.testing.assert_allclose(got.numpy(), want, rtol=1e-5, atol=1e-3)
    info = torch.tensor([[224., 224., 1.], [608., 800., 1.]])
    c = clip_boxes(got.clone(), info, 2).numpy()
    assert c[0].min() >= 0 and c[0].max() <= 223 and c[1, :, 0::2].max() <= 799 and c[1, :, 1::2].max() <= 607
