"""Pins the C restatement (oracle/nafae_oracle.c) to outputs of the reference's own, unmodified
CUDA kernels executed on a B200 (tests/golden/ref_gpu_*.npz, made by make_ref_gpu_golden.py)."""
import glob
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from _golden import GOLDEN
from oracle import cpu as ocpu

ROI_FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "ref_gpu_roi_*.npz")))


def _id(p):
    return os.path.basename(p)[len("ref_gpu_roi_"):-4]


def test_fixtures_present():
    assert ROI_FIXTURES and os.path.exists(os.path.join(GOLDEN, "ref_gpu_nms.npz"))


@pytest.mark.parametrize("path", ROI_FIXTURES, ids=_id)
def test_roi_align_forward_bit_exact(path):
    z = np.load(path)
    y = ocpu.roi_align_forward(z["features"], z["rois"], int(z["ah"]), int(z["aw"]),
                               float(z["scale"]))
    np.testing.assert_array_equal(y, z["align_fwd"])  # bit-for-bit, incl. the fp64 partial sums


@pytest.mark.parametrize("path", ROI_FIXTURES, ids=_id)
def test_roi_align_avg_max_modules_bit_exact(path):
    z = np.load(path)
    if "align_avg_fwd" not in z:
        pytest.skip("no pooled outputs")
    ah, aw, s = int(z["ah"]) - 1, int(z["aw"]) - 1, float(z["scale"])
    np.testing.assert_array_equal(ocpu.roi_align_avg_forward(z["features"], z["rois"], ah, aw, s),
                                  z["align_avg_fwd"])
    np.testing.assert_array_equal(ocpu.roi_align_max_forward(z["features"], z["rois"], ah, aw, s),
                                  z["align_max_fwd"])


@pytest.mark.parametrize("path", ROI_FIXTURES, ids=_id)
def test_roi_align_backward_close(path):
    """The reference accumulates with atomicAdd in unspecified order: tolerance, not bits."""
    z = np.load(path)
    fs, s = z["features"].shape, float(z["scale"])
    g = ocpu.roi_align_backward(z["align_top_diff"], z["rois"], fs, s)
    ref = z["align_bwd"]
    np.testing.assert_allclose(g, ref, rtol=1e-5, atol=1e-5 * np.abs(ref).max())
    if "align_avg_bwd" in z:
        for mode, fn in (("avg", ocpu.roi_align_avg_backward), ("max", ocpu.roi_align_max_backward)):
            g = fn(z["pooled_top_diff"], z["features"], z["rois"], s)
            ref = z["align_%s_bwd" % mode]
            np.testing.assert_allclose(g, ref, rtol=1e-5, atol=1e-5 * np.abs(ref).max(),
                                       err_msg=mode)


@pytest.mark.parametrize("path", ROI_FIXTURES, ids=_id)
def test_roi_pool_forward_backward_bit_exact(path):
    z = np.load(path)
    ph, pw = z["pool_fwd"].shape[2:]
    y, am = ocpu.roi_pool_forward(z["features"], z["rois"], ph, pw, float(z["scale"]))
    np.testing.assert_array_equal(y, z["pool_fwd"])
    np.testing.assert_array_equal(am, z["pool_argmax"])
    g = ocpu.roi_pool_backward(z["pool_top_diff"], am, z["rois"], z["features"].shape,
                               float(z["scale"]))
    np.testing.assert_array_equal(g, z["pool_bwd"])  # gather form: deterministic order


def test_pool2x2_matches_aten_cpu():
    rs = np.random.RandomState(1)
    x = rs.standard_normal((3, 4, 8, 8)).astype(np.float32)
    x[0, 0, 2, 3] = np.nan
    x[1, 1] = 0.5  # ties
    t = torch.from_numpy(x)
    np.testing.assert_array_equal(ocpu.pool2x2_forward(x, False),
                                  F.avg_pool2d(t, 2, 1).numpy())
    np.testing.assert_array_equal(ocpu.pool2x2_forward(x, True), F.max_pool2d(t, 2, 1).numpy())
    gy = rs.standard_normal((3, 4, 7, 7)).astype(np.float32)
    for is_max, fn in ((False, F.avg_pool2d), (True, F.max_pool2d)):
        tt = torch.from_numpy(np.nan_to_num(x)).requires_grad_(True)
        fn(tt, 2, 1).backward(torch.from_numpy(gy))
        got = ocpu.pool2x2_backward(np.nan_to_num(x), gy, is_max)
        np.testing.assert_allclose(got, tt.grad.numpy(), rtol=1e-6, atol=1e-6)


def test_roi_align_avg_backward_is_the_adjoint_of_the_forward():
    """Size-independent property that ties the checker's two directions together: RoIAlignAvg is linear
    in the features, so <forward(x), g> == <x, backward(g)> (the GPU backward kernels are compared
    against this backward, the forward against reference-GPU fixtures)."""
    from nafae_b200 import synth
    rs = np.random.RandomState(5)
    F, C, H, W, k = 3, 8, 20, 26, 7
    feat = (synth.conv5_maps(rs, F, C, H, W) - 0.3).astype(np.float32)
    props, _ = synth.proposals(rs, F, k, H * 16, W * 16)
    rois = np.concatenate([np.repeat(np.arange(F, dtype=np.float32), k)[:, None], props.reshape(-1, 4)], 1)
    rois[3, 1:] = 0  # a zero-padded proposal row
    g = rs.randn(F * k, C, 7, 7).astype(np.float32)
    y = ocpu.roi_align_avg_forward(feat, rois, 7, 7, 1 / 16.)
    gx = ocpu.roi_align_avg_backward(g, feat, rois, 1 / 16.)
    lhs = float((y.astype(np.float64) * g).sum())
    rhs = float((feat.astype(np.float64) * gx).sum())
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), abs(rhs), 1.0), (lhs, rhs)
