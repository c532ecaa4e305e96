"""NMS / proposal-tail parity: CUDA path (through the C ABI) vs the oracle, bit-exact."""
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


# ------------------------------------------------------------------ CPU: oracle itself ----
def _naive_nms(dets, thr):
    """Textbook greedy NMS using the oracle's IoU (independent of the mask/sweep structure)."""
    n = len(dets)
    removed = np.zeros(n, bool)
    keep = []
    for i in range(n):
        if removed[i]:
            continue
        keep.append(i)
        for j in range(i + 1, n):
            if not removed[j] and ocpu.iou(dets[i, :4], dets[j, :4]) > thr:
                removed[j] = True
    return np.asarray(keep, np.int32)


@pytest.mark.parametrize("name", ["n300", "n64", "n65", "n1", "n129_t03", "edge", "grid_t05"])
def test_oracle_mask_sweep_equals_naive_greedy(name):
    dets, thr = _cases.nms_cases()[name]
    np.testing.assert_array_equal(ocpu.nms(dets, thr), _naive_nms(dets, thr))


def test_oracle_nms_empty():
    assert ocpu.nms(np.zeros((0, 5), np.float32), 0.7).shape == (0,)


def test_oracle_iou_is_the_compiled_reference_formula():
    # a pair where fusing the column-box area changes the last bit of the union
    rs = np.random.RandomState(5)
    diff = 0
    for _ in range(2000):
        a = np.sort(rs.uniform(0, 800, 4)).astype(np.float32)[[0, 1, 2, 3]]
        b = np.sort(rs.uniform(0, 800, 4)).astype(np.float32)
        a = np.array([a[0], a[1], a[2], a[3]], np.float32)
        got = np.float32(ocpu.iou(a, b))
        f = np.float32
        inter = max(f(f(min(a[2], b[2]) - max(a[0], b[0])) + f(1)), f(0)) * \
            max(f(f(min(a[3], b[3]) - max(a[1], b[1])) + f(1)), f(0))
        sa = f(f(a[2] - a[0] + f(1)) * f(a[3] - a[1] + f(1)))
        sb = f(f(b[2] - b[0] + f(1)) * f(b[3] - b[1] + f(1)))
        unfused = f(inter) / f(f(sa + sb) - f(inter))
        diff += int(got != unfused)
        assert abs(float(got) - float(unfused)) <= 2e-7 * max(1.0, abs(float(unfused)))
    assert diff > 0  # the FMA contraction is observable: an unfused oracle would be wrong


def test_oracle_proposal_tail_pads_and_labels_frames():
    rs = np.random.RandomState(3)
    p, s = synth.proposals(rs, 3, 40, 608, 800)
    rois, rsc, nk = ocpu.proposal_tail(p, s, 6000, 30, 0.3)
    for f in range(3):
        d = np.concatenate([p[f], s[f][:, None]], 1)
        keep = ocpu.nms(d, 0.3)[:30]
        assert nk[f] == len(keep)
        np.testing.assert_array_equal(rois[f, :, 0], np.full(30, f, np.float32))
        np.testing.assert_array_equal(rois[f, : len(keep), 1:], p[f][keep])
        np.testing.assert_array_equal(rsc[f, : len(keep)], s[f][keep])
        assert not rois[f, len(keep):, 1:].any() and not rsc[f, len(keep):].any()


def _fixture(name):
    path = os.path.join(GOLDEN, name)
    if not os.path.exists(path):
        pytest.skip("fixture %s not generated yet" % name)
    return np.load(path)


def test_oracle_nms_matches_reference_gpu_fixture():
    """Keep lists produced by the reference's own nms_cuda_compute on a B200."""
    z = _fixture("ref_gpu_nms.npz")
    names = sorted(k[:-6] for k in z.files if k.endswith("__dets"))
    assert names
    for name in names:
        keep = ocpu.nms(z[name + "__dets"], float(z[name + "__thresh"]))
        np.testing.assert_array_equal(keep, z[name + "__keep"], err_msg=name)


# ------------------------------------------------------------------------ GPU: product ----
def _dev():
    return torch.device("cuda:0")


@gpu
@pytest.mark.parametrize("name", sorted(_cases.nms_cases()))
def test_nms_matches_oracle(name):
    from nafae_b200.model.nms.nms_wrapper import nms
    dets, thr = _cases.nms_cases()[name]
    keep = nms(torch.from_numpy(dets).to(_dev()), thr)
    assert keep.dtype == torch.int32 and keep.dim() == 2 and keep.shape[1] == 1
    np.testing.assert_array_equal(keep.cpu().numpy().ravel(), ocpu.nms(dets, thr))


@gpu
def test_nms_matches_reference_gpu_fixture():
    from nafae_b200.model.nms.nms_wrapper import nms
    z = _fixture("ref_gpu_nms.npz")
    for name in sorted(k[:-6] for k in z.files if k.endswith("__dets")):
        keep = nms(torch.from_numpy(z[name + "__dets"]).to(_dev()), float(z[name + "__thresh"]))
        np.testing.assert_array_equal(keep.cpu().numpy().ravel(), z[name + "__keep"], err_msg=name)


@gpu
def test_nms_matches_live_reference_kernels():
    """Same inputs through the reference's unmodified CUDA (oracle/_ref) on this GPU."""
    from oracle import ref_gpu
    if not ref_gpu.available():
        pytest.skip("oracle/_ref not built")
    from nafae_b200.model.nms.nms_wrapper import nms
    rs = np.random.RandomState(77)
    for n in (2352, 1000, 300, 77):
        p, s = synth.proposals(rs, 1, n, 608, 800)
        d = torch.from_numpy(np.concatenate([p[0], s[0][:, None]], 1)).to(_dev())
        for thr in (0.7, 0.5, 0.3):
            a = nms(d, thr).view(-1).cpu().numpy()
            b = ref_gpu.nms(d, thr).cpu().numpy()
            np.testing.assert_array_equal(a, b)
            np.testing.assert_array_equal(a, ocpu.nms(d.cpu().numpy(), thr))


@gpu
def test_nms_empty_and_contract():
    from nafae_b200.model.nms.nms_wrapper import nms
    assert nms(torch.zeros((0, 5), device=_dev()), 0.7) == []
    with pytest.raises(NotImplementedError):
        nms(torch.zeros((4, 5)), 0.7)  # CPU tensor: no CPU path


@gpu
def test_nms_batched_frames_independent():
    from nafae_b200.model.nms.nms_wrapper import nms_batched
    rs = np.random.RandomState(8)
    p, s = synth.proposals(rs, 5, 700, 608, 800)
    d = np.concatenate([p, s[:, :, None]], 2)
    keep, num = nms_batched(torch.from_numpy(d).to(_dev()), 0.7)
    keep, num = keep.cpu().numpy(), num.cpu().numpy()
    for f in range(5):
        np.testing.assert_array_equal(keep[f, : num[f]], ocpu.nms(d[f], 0.7))


@gpu
def test_reference_named_entry_point_nms_cuda_compute():
    """The symbol the reference's cffi glue binds (nms_cuda.c:12-17), same argument meaning."""
    import ctypes
    from nafae_b200 import _C
    dets, thr = _cases.nms_cases()["n300"]
    d = torch.from_numpy(dets).to(_dev())
    keep = torch.zeros(len(dets), dtype=torch.int32, device=_dev())
    num = torch.zeros(1, dtype=torch.int32, device=_dev())
    torch.cuda.synchronize()
    _C.lib.nms_cuda_compute(_C.ptr(keep), _C.ptr(num), _C.ptr(d), len(dets), 5, ctypes.c_float(thr))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(keep[: int(num[0])].cpu().numpy(), ocpu.nms(dets, thr))


@gpu
@pytest.mark.parametrize("F,n,pre,post,thr", [(40, 2352, 6000, 20, 0.7), (32, 300, 6000, 100, 0.7),
                                              (3, 500, 100, 50, 0.5), (2, 10, 6000, 20, 0.7),
                                              (4, 1000, 6000, 300, 0.3), (1, 64, 0, 7, 0.7)])
def test_proposal_tail_matches_oracle(F, n, pre, post, thr):
    from nafae_b200.model.rpn.proposal_layer import proposal_tail
    rs = np.random.RandomState(F * 1000 + n)
    p, s = synth.proposals(rs, F, n, 608, 800)
    rois, rsc, nk = proposal_tail(torch.from_numpy(p).to(_dev()), torch.from_numpy(s).to(_dev()),
                                  pre, post, thr, return_num=True)
    o_rois, o_rsc, o_nk = ocpu.proposal_tail(p, s, pre, post, thr)
    np.testing.assert_array_equal(nk.cpu().numpy(), o_nk)
    np.testing.assert_array_equal(rois.cpu().numpy(), o_rois)
    np.testing.assert_array_equal(rsc.cpu().numpy(), o_rsc)


@gpu
def test_proposal_tail_equals_reference_python_loop():
    """Restates proposal_layer.py:130-163 with this package's nms() as the reference would run it."""
    from nafae_b200.model.nms.nms_wrapper import nms
    from nafae_b200.model.rpn.proposal_layer import proposal_tail
    rs = np.random.RandomState(99)
    F, n, post = 6, 900, 20
    p, s = synth.proposals(rs, F, n, 608, 800)
    pt, st = torch.from_numpy(p).to(_dev()), torch.from_numpy(s).to(_dev())
    out = st.new_zeros(F, post, 5)
    out_s = st.new_zeros(F, post)
    for i in range(F):
        keep = nms(torch.cat((pt[i], st[i].view(-1, 1)), 1), 0.7).long().view(-1)[:post]
        out[i, :, 0] = i
        out[i, : keep.numel(), 1:] = pt[i][keep]
        out_s[i, : keep.numel()] = st[i][keep]
    rois, rsc = proposal_tail(pt, st, 6000, post, 0.7)
    assert torch.equal(rois, out) and torch.equal(rsc, out_s)


@gpu
def test_proposal_layer_module_equals_reference_style_loop():
    """`_ProposalLayer` mirror (anchors + decode + clip + sort + fused tail) vs the reference's own
    per-frame loop (proposal_layer.py:130-163) run with this package's nms()."""
    import types
    from nafae_b200.model.nms.nms_wrapper import nms
    from nafae_b200.model.rpn.proposal_layer import _ProposalLayer
    from nafae_b200.model.rpn.bbox_transform import bbox_transform_inv, clip_boxes
    cfg = types.SimpleNamespace(TEST=types.SimpleNamespace(RPN_PRE_NMS_TOP_N=6000, RPN_POST_NMS_TOP_N=20,
                                                          RPN_NMS_THRESH=0.7, RPN_MIN_SIZE=16))
    layer = _ProposalLayer(16, [4, 8, 16, 32], [0.5, 1, 2], cfg)
    torch.manual_seed(3)
    B, A, H, W = 4, 12, 14, 14
    cls = torch.rand(B, 2 * A, H, W, device=_dev())
    deltas = torch.randn(B, 4 * A, H, W, device=_dev()) * 0.2
    info = torch.tensor([[224., 224., 1.]] * B, device=_dev())
    out = layer((cls, deltas, info, "TEST"))
    assert out.shape == (B, 20, 5)
    # reference-style loop on the same decoded boxes
    scores = cls[:, A:].permute(0, 2, 3, 1).contiguous().view(B, -1)
    sx = torch.arange(0, W, device=_dev(), dtype=torch.float32) * 16
    sy = torch.arange(0, H, device=_dev(), dtype=torch.float32) * 16
    yy, xx = torch.meshgrid(sy, sx, indexing="ij")
    shifts = torch.stack((xx.reshape(-1), yy.reshape(-1), xx.reshape(-1), yy.reshape(-1)), 1)
    anchors = (layer._anchors.to(_dev()).view(1, A, 4) + shifts.view(-1, 1, 4)).view(1, -1, 4).expand(B, -1, 4)
    props = clip_boxes(bbox_transform_inv(anchors, deltas.permute(0, 2, 3, 1).contiguous().view(B, -1, 4), B),
                       info, B)
    _, order = torch.sort(scores, 1, True)
    want = scores.new_zeros(B, 20, 5)
    want_s = scores.new_zeros(B, 20)
    for i in range(B):
        p, s = props[i][order[i]], scores[i][order[i]].view(-1, 1)
        keep = nms(torch.cat((p, s), 1), 0.7).long().view(-1)[:20]
        want[i, :, 0] = i
        want[i, : keep.numel(), 1:] = p[keep]
        want_s[i, : keep.numel()] = s[keep, 0]
    assert torch.equal(out, want) and torch.equal(layer.get_roi_score(), want_s)
