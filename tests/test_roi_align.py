"""RoIAlign / RoIAlignAvg / RoIAlignMax parity: CUDA path vs oracle and reference-GPU fixtures."""
import glob
import os

import numpy as np
import pytest
import torch

import _cases
from _golden import GOLDEN
from nafae_b200 import synth
from oracle import cpu as ocpu

gpu = pytest.mark.gpu
ROI_FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "ref_gpu_roi_*.npz")))
# tolerance of the fp32 ("fast") path, north_star: <= 1e-4 relative for fp32 pooled features
RTOL = 1e-4


def _id(p):
    return os.path.basename(p)[len("ref_gpu_roi_"):-4]


def _dev():
    return torch.device("cuda:0")


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(_dev())


def _mods():
    from nafae_b200.model.roi_align.modules.roi_align import RoIAlign, RoIAlignAvg, RoIAlignMax
    return RoIAlign, RoIAlignAvg, RoIAlignMax


def _close(got, ref, rtol=RTOL):
    scale = float(np.abs(ref).max()) if ref.size else 1.0
    np.testing.assert_allclose(got, ref, rtol=rtol, atol=rtol * 1e-1 * max(scale, 1e-30))


@gpu
@pytest.mark.parametrize("path", ROI_FIXTURES, ids=_id)
def test_exact_mode_bit_identical_to_reference_gpu_fixture(path):
    RoIAlign, RoIAlignAvg, RoIAlignMax = _mods()
    z = np.load(path)
    ah, aw, s = int(z["ah"]), int(z["aw"]), float(z["scale"])
    f, r = _t(z["features"]), _t(z["rois"])
    y = RoIAlign(ah, aw, s, exact=True)(f, r)
    np.testing.assert_array_equal(y.cpu().numpy(), z["align_fwd"])
    np.testing.assert_array_equal(RoIAlignAvg(ah - 1, aw - 1, s, exact=True)(f, r).cpu().numpy(),
                                  z["align_avg_fwd"])
    np.testing.assert_array_equal(RoIAlignMax(ah - 1, aw - 1, s, exact=True)(f, r).cpu().numpy(),
                                  z["align_max_fwd"])


@gpu
@pytest.mark.parametrize("path", ROI_FIXTURES, ids=_id)
def test_fast_mode_within_tolerance_of_fixture(path):
    RoIAlign, RoIAlignAvg, RoIAlignMax = _mods()
    z = np.load(path)
    ah, aw, s = int(z["ah"]), int(z["aw"]), float(z["scale"])
    f, r = _t(z["features"]), _t(z["rois"])
    _close(RoIAlign(ah, aw, s)(f, r).cpu().numpy(), z["align_fwd"])
    _close(RoIAlignAvg(ah - 1, aw - 1, s)(f, r).cpu().numpy(), z["align_avg_fwd"])
    _close(RoIAlignMax(ah - 1, aw - 1, s)(f, r).cpu().numpy(), z["align_max_fwd"])


@gpu
@pytest.mark.parametrize("path", ROI_FIXTURES, ids=_id)
@pytest.mark.parametrize("exact", [True, False])
def test_backward_matches_fixture(path, exact):
    RoIAlign, RoIAlignAvg, RoIAlignMax = _mods()
    z = np.load(path)
    ah, aw, s = int(z["ah"]), int(z["aw"]), float(z["scale"])
    r = _t(z["rois"])
    for mod, td, ref in ((RoIAlign(ah, aw, s, exact=exact), z["align_top_diff"], z["align_bwd"]),
                         (RoIAlignAvg(ah - 1, aw - 1, s, exact=exact), z["pooled_top_diff"],
                          z["align_avg_bwd"]),
                         (RoIAlignMax(ah - 1, aw - 1, s, exact=exact), z["pooled_top_diff"],
                          z["align_max_bwd"])):
        f = _t(z["features"]).requires_grad_(True)
        mod(f, r).backward(_t(td))
        # atomics: order is unspecified in the reference too -> tolerance
        _close(f.grad.cpu().numpy(), ref, rtol=1e-5 if exact else RTOL)


@gpu
def test_live_reference_kernels_bit_exact():
    """Exact mode vs the reference's unmodified CUDA on the same GPU, benchmark-shaped inputs."""
    from oracle import ref_gpu
    if not ref_gpu.available():
        pytest.skip("oracle/_ref not built")
    RoIAlign, RoIAlignAvg, _ = _mods()
    rs = np.random.RandomState(5)
    F, C, H, W, Nb = 3, 32, 38, 50, 20
    feat = _t(synth.conv5_maps(rs, F, C, H, W))
    p, _ = synth.proposals(rs, F, Nb, 608, 800)
    rois = np.concatenate([np.repeat(np.arange(F, dtype=np.float32), Nb)[:, None],
                           p.reshape(-1, 4)], 1)
    rois[7, 1:] = 0  # a zero-padded proposal row
    r = _t(rois)
    ref = ref_gpu.roi_align_forward(feat, r, 8, 8, 1 / 16.)
    assert torch.equal(RoIAlign(8, 8, 1 / 16., exact=True)(feat, r), ref)
    ref_avg = torch.nn.functional.avg_pool2d(ref, kernel_size=2, stride=1)
    assert torch.equal(RoIAlignAvg(7, 7, 1 / 16., exact=True)(feat, r), ref_avg)
    fast = RoIAlignAvg(7, 7, 1 / 16.)(feat, r)  # slab kernel
    _close(fast.cpu().numpy(), ref_avg.cpu().numpy())


def _frame_rois(rs, F, per_frame, img_h, img_w, shuffle=False):
    rows = []
    for f, k in enumerate(per_frame):
        if k == 0:
            continue
        p, _ = synth.proposals(rs, 1, k, img_h, img_w)
        rows.append(np.concatenate([np.full((k, 1), f, np.float32), p[0]], 1))
    rois = np.concatenate(rows, 0)
    if shuffle:
        rois = rois[rs.permutation(len(rois))]
    return rois


@gpu
@pytest.mark.parametrize("name,F,C,H,W,per_frame,shuffle", [
    ("cfg2_like", 6, 64, 38, 50, [20] * 6, False),
    ("real_14x14", 5, 64, 14, 14, [20] * 5, False),
    ("ragged_shuffled", 5, 32, 38, 50, [3, 0, 40, 1, 17], True),
    ("chunked_300_in_one_frame", 3, 8, 38, 50, [300, 5, 130], True),
    ("many_frames_few_channels", 40, 8, 14, 14, [4] * 40, False),
    ("odd_map_generic_fallback", 2, 8, 37, 50, [9, 9], False),
    ("c_not_multiple_of_4", 2, 6, 38, 50, [9, 9], False),
])
@pytest.mark.parametrize("pool", ["avg", "max"])
def test_slab_kernel_matches_oracle(name, F, C, H, W, per_frame, shuffle, pool):
    _, RoIAlignAvg, RoIAlignMax = _mods()
    rs = np.random.RandomState(abs(hash(name)) % 1000)
    feat = synth.conv5_maps(rs, F, C, H, W) - 0.3  # signed values
    rois = _frame_rois(rs, F, per_frame, H * 16, W * 16, shuffle)
    mod = (RoIAlignAvg if pool == "avg" else RoIAlignMax)(7, 7, 1 / 16.)
    got = mod(_t(feat), _t(rois)).cpu().numpy()
    ofn = ocpu.roi_align_avg_forward if pool == "avg" else ocpu.roi_align_max_forward
    _close(got, ofn(feat, rois, 7, 7, 1 / 16.))


@gpu
@pytest.mark.parametrize("name,F,C,H,W,per_frame,shuffle", [
    ("cfg2_like", 6, 64, 38, 50, [20] * 6, False),
    ("real_14x14", 5, 64, 14, 14, [20] * 5, False),
    ("ragged_shuffled_empty_frame", 5, 32, 38, 50, [3, 0, 40, 1, 17], True),
    ("chunked_300_in_one_frame", 3, 8, 38, 50, [300, 5, 130], True),
    ("many_frames_few_channels", 40, 8, 14, 14, [4] * 40, False),
    ("odd_map_generic_fallback", 2, 8, 37, 50, [9, 9], False),
])
def test_avg_backward_through_module_matches_oracle(name, F, C, H, W, per_frame, shuffle):
    """RoIAlignAvg backward through the module at slab-kernel shapes (ragged / empty frames,
    >table-size frames, fallback shapes): frames without RoIs must come back as exact zeros."""
    _, RoIAlignAvg, _ = _mods()
    rs = np.random.RandomState(abs(hash(name)) % 1000)
    feat = synth.conv5_maps(rs, F, C, H, W)
    rois = _frame_rois(rs, F, per_frame, H * 16, W * 16, shuffle)
    gy = rs.randn(rois.shape[0], C, 7, 7).astype(np.float32)
    f = _t(feat).requires_grad_(True)
    junk = torch.full((F * C * H * W * 2,), float("nan"), device=_dev())  # poison the allocator
    del junk
    RoIAlignAvg(7, 7, 1 / 16.)(f, _t(rois)).backward(_t(gy))
    ref = ocpu.roi_align_avg_backward(gy, feat, rois, 1 / 16.)
    got = f.grad.cpu().numpy()
    assert np.isfinite(got).all()
    np.testing.assert_allclose(got, ref, rtol=RTOL, atol=RTOL * np.abs(ref).max())


@gpu
@pytest.mark.parametrize("name,F,C,H,W,per_frame,shuffle", [
    ("cfg2_like", 4, 32, 38, 50, [20] * 4, False),
    ("ragged_shuffled_empty_frame", 5, 16, 38, 50, [3, 0, 40, 1, 17], True),
    ("more_than_one_id_table", 2, 8, 14, 14, [300, 131], True),
    ("odd_width", 2, 8, 20, 37, [9, 12], False),
])
def test_avg_backward_kernels_through_the_c_abi(name, F, C, H, W, per_frame, shuffle):
    """nafae_roi_align_backward for RoIAlignAvg 7x7: the shared-memory scatter overwriting and
    accumulating (cp.reduce bulk add onto a non-zero tensor), the deterministic cell-gather (bitwise
    reproducible) and the reference-style global-atomic kernel all agree with the oracle."""
    from nafae_b200 import _C
    rs = np.random.RandomState(abs(hash(name)) % 1000)
    feat_shape = (F, C, H, W)
    rois = _frame_rois(rs, F, per_frame, H * 16, W * 16, shuffle)
    gy = rs.randn(rois.shape[0], C, 7, 7).astype(np.float32)
    ref = ocpu.roi_align_avg_backward(gy, np.zeros(feat_shape, np.float32), rois, 1 / 16.)
    tol = dict(rtol=RTOL, atol=RTOL * np.abs(ref).max())
    g, r = _t(gy), _t(rois)

    def call(out, flags):
        st = _C.lib.nafae_roi_align_backward(_C.ptr(g), None, 1 / 16., F, rois.shape[0], H, W, C, 7, 7, _C.POOL_AVG,
                                             _C.ptr(r), _C.ptr(out), flags, _C.stream())
        assert st == 1
        torch.cuda.synchronize()
        return out.cpu().numpy()

    nan = lambda: torch.full(feat_shape, float("nan"), device=_dev())
    np.testing.assert_allclose(call(nan(), _C.FLAG_OVERWRITE), ref, **tol)
    base = rs.randn(*feat_shape).astype(np.float32)
    np.testing.assert_allclose(call(_t(base.copy()), 0), base + ref, **tol)                      # accumulate
    np.testing.assert_allclose(call(_t(base.copy()), _C.FLAG_EXACT), base + ref, **tol)          # global atomics
    det = call(nan(), _C.FLAG_OVERWRITE | _C.FLAG_DETERMINISTIC)
    np.testing.assert_allclose(det, ref, **tol)
    assert np.array_equal(det, call(nan(), _C.FLAG_OVERWRITE | _C.FLAG_DETERMINISTIC))


@gpu
def test_out_of_range_batch_index_rows_are_zero():
    _, RoIAlignAvg, _ = _mods()
    rs = np.random.RandomState(2)
    feat = synth.conv5_maps(rs, 2, 8, 38, 50)
    rois = _frame_rois(rs, 2, [4, 4], 608, 800)
    rois[2, 0] = 7
    rois[5, 0] = -1
    got = RoIAlignAvg(7, 7, 1 / 16.)(_t(feat), _t(rois)).cpu().numpy()
    assert not got[2].any() and not got[5].any()
    ok = [0, 1, 3, 4, 6, 7]
    _close(got[ok], ocpu.roi_align_avg_forward(feat, rois[ok], 7, 7, 1 / 16.))


@gpu
def test_cfg2_full_size_against_oracle_and_linearity():
    """BASELINE cfg2 shapes: 40 x 512 x 38 x 50 maps, 800 RoIs -> (800, 512, 7, 7)."""
    _, RoIAlignAvg, _ = _mods()
    c = synth.CONFIGS["cfg2"]
    rs = np.random.RandomState(1234 + 2)
    F = c["Na"] * c["Ns"]
    feat = synth.conv5_maps(rs, F, c["C"], c["H"], c["W"])
    rois = _frame_rois(rs, F, [c["Nb"]] * F, c["img_h"], c["img_w"])
    rois[13, 1:] = 0
    mod = RoIAlignAvg(7, 7, 1 / 16.)
    ft, rt = _t(feat), _t(rois)
    got = mod(ft, rt)
    ref = ocpu.roi_align_avg_forward(feat, rois, 7, 7, 1 / 16.)
    _close(got.cpu().numpy(), ref)
    # size-independent property: the operator is linear in the features
    other = torch.randn_like(ft)
    lhs = mod(ft + 2 * other, rt)
    rhs = got + 2 * mod(other, rt)
    assert torch.allclose(lhs, rhs, rtol=1e-4, atol=1e-4)


@gpu
def test_module_contract():
    RoIAlign, RoIAlignAvg, _ = _mods()
    with pytest.raises(NotImplementedError):
        RoIAlignAvg(7, 7, 1 / 16.)(torch.zeros(1, 4, 8, 8), torch.zeros(1, 5))  # CPU tensors
    f = torch.zeros(1, 4, 8, 8, device=_dev())
    with pytest.raises(ValueError):
        RoIAlign(7, 7, 1 / 16.)(f, torch.zeros(3, 4, device=_dev()))  # size_rois != 5
    out = RoIAlignAvg(7, 7, 1 / 16.)(f, torch.zeros(0, 5, device=_dev()))
    assert out.shape == (0, 4, 7, 7)
    # differentiable w.r.t. features only, backward returns (grad, None)
    f = torch.randn(2, 4, 9, 11, device=_dev(), requires_grad=True)
    r = torch.tensor([[0, 0, 0, 60, 40], [1, 8, 8, 100, 90.]], device=_dev(), requires_grad=True)
    RoIAlignAvg(7, 7, 1 / 16.)(f, r).sum().backward()
    assert f.grad is not None and r.grad is None


@gpu
def test_reference_named_launchers():
    """ROIAlignForwardLaucher / ROIAlignBackwardLaucher: the symbols the reference glue binds."""
    from nafae_b200 import _C
    c = _cases.roi_cases()["small_8x8"]
    f, r = _t(c["features"]), _t(c["rois"])
    B, C, H, W = f.shape
    R = r.shape[0]
    out = torch.zeros((R, C, 8, 8), device=_dev())
    st = _C.lib.ROIAlignForwardLaucher(_C.ptr(f), c["scale"], R, H, W, C, 8, 8, _C.ptr(r),
                                       _C.ptr(out), _C.stream())
    assert st == 1
    np.testing.assert_array_equal(out.cpu().numpy(),
                                  ocpu.roi_align_forward(c["features"], c["rois"], 8, 8, c["scale"]))
    td = _t(_cases.top_diff_for(c, 8, 8))
    bd = torch.zeros_like(f)
    st = _C.lib.ROIAlignBackwardLaucher(_C.ptr(td), c["scale"], B, R, H, W, C, 8, 8, _C.ptr(r),
                                        _C.ptr(bd), _C.stream())
    assert st == 1
    ref = ocpu.roi_align_backward(td.cpu().numpy(), c["rois"], f.shape, c["scale"])
    _close(bd.cpu().numpy(), ref, rtol=1e-5)


@gpu
@pytest.mark.parametrize("name,F,C,H,W,per_frame,shuffle", [
    ("cfg2_like", 40, 64, 38, 50, [20] * 40, False),
    ("real_14x14", 40, 64, 14, 14, [20] * 40, False),
    ("ragged_shuffled_empty_frames", 9, 32, 38, 50, [3, 0, 40, 1, 17, 0, 60, 52, 53], True),
    ("whole_table_and_chunked", 4, 32, 38, 50, [300, 5, 130, 104], True),
    ("more_than_1024_rois", 12, 8, 38, 50, [100] * 12, False),
])
def test_persistent_grid_size_and_gate_workspace_do_not_change_results(name, F, C, H, W, per_frame, shuffle):
    """The persistent kernel splits (frame, channel-group) units over however many SMs it is given:
    148, a handful, or one CTA must produce BIT-identical pooled features, with or without the
    residency-gate workspace, launch after launch; the gate counts exactly the gated launches."""
    from nafae_b200 import _C
    rs = np.random.RandomState(len(name) * 7 + F)
    feat = _t(synth.conv5_maps(rs, F, C, H, W) - 0.3)
    rois_np = _frame_rois(rs, F, per_frame, H * 16, W * 16, shuffle)
    rois = _t(rois_np)
    R = rois.shape[0]

    def run(ws, flags=0):
        out = torch.full((R, C, 7, 7), float("nan"), device=_dev())
        st = _C.lib.nafae_roi_align_forward(_C.ptr(feat), 1 / 16., F, R, H, W, C, 7, 7, _C.POOL_AVG,
                                            _C.ptr(rois), _C.ptr(out), flags, _C.ptr(ws),
                                            ws.numel() * 4 if ws is not None else 0, _C.stream())
        assert st == 1, _C.last_error()
        return out
    base = run(None)
    _close(base.cpu().numpy(), ocpu.roi_align_avg_forward(feat.cpu().numpy(), rois_np, 7, 7, 1 / 16.))
    nbytes = int(_C.lib.nafae_roi_align_workspace_bytes(F, R))
    assert nbytes >= _C.ROI_ALIGN_WS_BYTES
    ws = torch.zeros(nbytes // 4, dtype=torch.int32, device=_dev())
    for it in range(3):
        got = run(ws, _C.FLAG_NO_GATE if it == 1 else 0)
        torch.cuda.synchronize()
        assert torch.equal(got, base), (name, it)
        assert int(ws[0]) == 0
    assert int(ws[1]) == 2  # two gated launches opened the gate, the NO_GATE one did not
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    prev = _C.lib.nafae_set_reserved_sms(0)
    try:
        for ctas in (8, 3, 2, 1):
            _C.lib.nafae_set_reserved_sms(sms - ctas)
            assert torch.equal(run(ws), base), (name, ctas)
            assert torch.equal(run(None), base), (name, ctas, "no workspace")
    finally:
        _C.lib.nafae_set_reserved_sms(prev)


@gpu
@pytest.mark.parametrize("F,C,H,W,Nb", [(6, 64, 38, 50, 20), (5, 64, 14, 14, 20)])
def test_bf16_output_is_the_rounded_fp32_output(F, C, H, W, Nb):
    """NAFAE_FLAG_OUT_BF16: the bandwidth kernel writes the bridge GEMM's A operand directly --
    (R, C*7*7) row-major bf16 -- and it is exactly the fp32 result rounded to nearest."""
    from nafae_b200 import _C
    rs = np.random.RandomState(F + C)
    feat = _t(synth.conv5_maps(rs, F, C, H, W) - 0.3)
    rois = _t(_frame_rois(rs, F, [Nb] * F, H * 16, W * 16))
    R = rois.shape[0]
    outs = []
    for flags, dt in ((0, torch.float32), (_C.FLAG_OUT_BF16, torch.bfloat16)):
        out = torch.full((R, C, 7, 7), float("nan"), dtype=dt, device=_dev())
        st = _C.lib.nafae_roi_align_forward(_C.ptr(feat), 1 / 16., F, R, H, W, C, 7, 7, _C.POOL_AVG, _C.ptr(rois),
                                            _C.ptr(out), flags, None, 0, _C.stream())
        assert st == 1, _C.last_error()
        outs.append(out)
    torch.cuda.synchronize()
    assert torch.equal(outs[1], outs[0].to(torch.bfloat16))
    # shapes the bandwidth kernel cannot take have no bf16 form: an argument error, not a silent cast
    odd = torch.zeros((2, 6, 37, 50), device=_dev())
    out = torch.zeros((R, 6, 7, 7), dtype=torch.bfloat16, device=_dev())
    assert _C.lib.nafae_roi_align_forward(_C.ptr(odd), 1 / 16., 2, R, 37, 50, 6, 7, 7, _C.POOL_AVG, _C.ptr(rois),
                                          _C.ptr(out), _C.FLAG_OUT_BF16, None, 0, _C.stream()) == 0
