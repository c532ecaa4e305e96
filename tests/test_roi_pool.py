"""RoIPool parity (bit-exact: compare / index work only)."""
import glob
import os

import numpy as np
import pytest
import torch

from _golden import GOLDEN
from nafae_b200 import synth
from oracle import cpu as ocpu

gpu = pytest.mark.gpu
ROI_FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "ref_gpu_roi_*.npz")))


def _id(p):
    return os.path.basename(p)[len("ref_gpu_roi_"):-4]


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda:0")


@gpu
@pytest.mark.parametrize("path", ROI_FIXTURES, ids=_id)
def test_roi_pool_matches_reference_gpu_fixture(path):
    from nafae_b200.model.roi_pooling.functions.roi_pool import RoIPoolFunction
    z = np.load(path)
    ph, pw = z["pool_fwd"].shape[2:]
    f = _t(z["features"]).requires_grad_(True)
    fn = RoIPoolFunction(ph, pw, float(z["scale"]))
    y = fn(f, _t(z["rois"]))
    np.testing.assert_array_equal(y.detach().cpu().numpy(), z["pool_fwd"])
    np.testing.assert_array_equal(fn.argmax.cpu().numpy(), z["pool_argmax"])
    y.backward(_t(z["pool_top_diff"]))
    np.testing.assert_array_equal(f.grad.cpu().numpy(), z["pool_bwd"])


@gpu
@pytest.mark.parametrize("F,C,H,W,k", [(4, 16, 38, 50, 20), (3, 8, 14, 14, 20), (2, 4, 38, 50, 300)])
def test_roi_pool_module_matches_oracle(F, C, H, W, k):
    from nafae_b200.model.roi_pooling.modules.roi_pool import _RoIPooling
    rs = np.random.RandomState(F * 100 + C)
    feat = synth.conv5_maps(rs, F, C, H, W) - 0.2
    p, _ = synth.proposals(rs, F, k, H * 16, W * 16)
    rois = np.concatenate([np.repeat(np.arange(F, dtype=np.float32), k)[:, None],
                           p.reshape(-1, 4)], 1)
    rois = rois[rs.permutation(len(rois))]
    rois[0, 1:] = 0
    f = _t(feat).requires_grad_(True)
    y = _RoIPooling(7, 7, 1 / 16.)(f, _t(rois))
    oy, oam = ocpu.roi_pool_forward(feat, rois, 7, 7, 1 / 16.)
    np.testing.assert_array_equal(y.detach().cpu().numpy(), oy)
    td = rs.standard_normal(oy.shape).astype(np.float32)
    y.backward(_t(td))
    og = ocpu.roi_pool_backward(td, oam, rois, feat.shape, 1 / 16.)
    np.testing.assert_array_equal(f.grad.cpu().numpy(), og)


@gpu
def test_roi_pool_live_reference():
    from oracle import ref_gpu
    if not ref_gpu.available():
        pytest.skip("oracle/_ref not built")
    from nafae_b200.model.roi_pooling.functions.roi_pool import RoIPoolFunction
    rs = np.random.RandomState(9)
    F, C, H, W, k = 3, 32, 38, 50, 20
    feat = _t(synth.conv5_maps(rs, F, C, H, W))
    p, _ = synth.proposals(rs, F, k, 608, 800)
    rois = _t(np.concatenate([np.repeat(np.arange(F, dtype=np.float32), k)[:, None],
                              p.reshape(-1, 4)], 1))
    fn = RoIPoolFunction(7, 7, 1 / 16.)
    f = feat.clone().requires_grad_(True)
    y = fn(f, rois)
    ry, ram = ref_gpu.roi_pool_forward(feat, rois, 7, 7, 1 / 16.)
    assert torch.equal(y.detach(), ry) and torch.equal(fn.argmax, ram)
    td = torch.randn_like(ry)
    y.backward(td)
    assert torch.equal(f.grad, ref_gpu.roi_pool_backward(td, ram, rois, feat.shape, 1 / 16.))
