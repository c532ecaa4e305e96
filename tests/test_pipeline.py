"""Whole-path parity at BASELINE.json's sizes: GroundingStep (the C-ABI calls bench.py times) vs the
oracle, graph replay == eager, pipelined schedule == sequential, cfg4 dense stress, cfg5 sweep."""
import numpy as np
import pytest
import torch

from nafae_b200 import synth
from oracle import cpu as ocpu
from oracle import dvsa as odvsa

gpu = pytest.mark.gpu
RTOL = 1e-4


def _step(cfg, seed, dev="cuda:0", tensor_cores=False):
    from nafae_b200.pipeline import GroundingStep
    c = synth.CONFIGS[cfg]
    st = GroundingStep(c["Na"], c["Ns"], c["Nb"], c["Ne"], c["D"], c["C"], c["H"], c["W"], c["n"],
                       pre_nms_topn=c["pre"], Delta=c["Delta"], vis_lam=c["vis_lam"], train=c["train"],
                       device=dev, tensor_cores=tensor_cores)
    b = synth.make_batch(cfg, seed)
    st.load(b)
    return c, b, st


def _check_against_oracle(c, b, st, check_pooled=True):
    rois, rsc, _ = ocpu.proposal_tail(b["proposals"], b["scores"], c["pre"], c["Nb"], 0.7)
    np.testing.assert_array_equal(st.rois.cpu().numpy(), rois)        # bit-exact keeps + padding
    np.testing.assert_array_equal(st.roi_scores.cpu().numpy(), rsc)
    if check_pooled:
        pooled = ocpu.roi_align_avg_forward(b["features"], rois.reshape(-1, 5), 7, 7, 1 / 16.)
        np.testing.assert_allclose(st.pooled.cpu().numpy(), pooled, rtol=RTOL, atol=1e-5)
    phase = "train" if c["train"] else "eval"
    ref = odvsa.dvsa_forward_backward(b["vis_feats"], b["word_feats"], b["lens"], c["Na"], c["Nb"],
                                      c["Ne"], c["Delta"], c["vis_lam"], phase)
    live = np.zeros((c["Na"], c["Ne"]), bool)
    for a, n in enumerate(b["lens"]):
        live[a, :n] = True
    live = live.reshape(-1)
    np.testing.assert_array_equal(st.D_ind.cpu().numpy()[:, live], ref["D_ind"][:, live])
    np.testing.assert_allclose(st.D_sim.cpu().numpy(), ref["D_sim"], rtol=RTOL, atol=1e-5)
    np.testing.assert_allclose(float(st.loss), ref["margin_loss"], rtol=RTOL)
    return ref


HEADS = pytest.mark.parametrize("tensor_cores", [False, True], ids=["fma_head", "tcgen05_head"])


@gpu
@HEADS
def test_cfg2_step_matches_oracle_and_graph_replay_is_identical(tensor_cores):
    c, b, st = _step("cfg2", 1234, tensor_cores=tensor_cores)
    st.run()
    torch.cuda.synchronize()
    ref = _check_against_oracle(c, b, st)
    for got, want in ((st.grad_vis, ref["grad_vis"]), (st.grad_word, ref["grad_word"])):
        np.testing.assert_allclose(got.cpu().numpy(), want, rtol=RTOL, atol=1e-5 * np.abs(want).max())
    eager = [t.clone() for t in (st.rois, st.pooled, st.D_ind, st.D_sim, st.loss, st.grad_word)]
    gv = st.grad_vis.clone()
    st.capture()
    for _ in range(3):
        st.replay()
    torch.cuda.synchronize()
    for e, t in zip(eager, (st.rois, st.pooled, st.D_ind, st.D_sim, st.loss, st.grad_word)):
        assert torch.equal(e, t)
    # the clustering gradient is accumulated with atomics: order may differ between runs
    assert torch.allclose(gv, st.grad_vis, rtol=1e-5, atol=1e-7)


@gpu
@HEADS
def test_cfg4_dense_stress_zero_padding_and_3200_rois(tensor_cores):
    """300 proposals/frame -> NMS 0.7 -> top-100 (fewer survive: zero-padded rows) -> RoIAlign."""
    c, b, st = _step("cfg4", 77, tensor_cores=tensor_cores)
    st.run()
    torch.cuda.synchronize()
    rois = st.rois.cpu().numpy()
    assert (rois[:, :, 1:].reshape(-1, 4).sum(1) == 0).any()  # padding rows exist
    _check_against_oracle(c, b, st)


@gpu
def test_cfg2_real_14x14_maps():
    c, b, st = _step("cfg2_real", 5)
    st.run()
    torch.cuda.synchronize()
    _check_against_oracle(c, b, st)


@gpu
@HEADS
def test_pipelined_schedule_equals_sequential(tensor_cores):
    """bench.py's three-branch graphs reorder work across batches, never inside one."""
    from nafae_b200 import _C
    from nafae_b200.pipeline import capture_pipelined
    c, b0, s0 = _step("cfg2", 1, tensor_cores=tensor_cores)
    _, b1, s1 = _step("cfg2", 2, tensor_cores=tensor_cores)
    steps = [s0, s1]
    for st in steps:
        st.run()
    torch.cuda.synchronize()
    want = [[t.clone() for t in (st.rois, st.pooled, st.D_ind, st.loss, st.grad_word)] for st in steps]
    prev = _C.lib.nafae_set_reserved_sms(16)
    try:
        side = [torch.cuda.Stream(), torch.cuda.Stream()]
        graphs = [capture_pipelined(steps[j], steps[1 - j], side) for j in range(2)]
        for st in steps:  # scribble over the outputs so that the replays must recompute them
            st.pooled.zero_(); st.rois.zero_(); st.D_ind.zero_(); st.grad_word.zero_(); st.loss.zero_()
        for i in range(6):
            graphs[i & 1].replay()
        torch.cuda.synchronize()
    finally:
        _C.lib.nafae_set_reserved_sms(prev)
    for st, w in zip(steps, want):
        for a, t in zip(w, (st.rois, st.pooled, st.D_ind, st.loss, st.grad_word)):
            assert torch.equal(a, t)


@gpu
@HEADS
def test_dependency_exact_multi_step_graph_equals_sequential(tensor_cores):
    """capture_pipelined_exact: 4 steps per graph with only the true dependencies between them; three
    replays (12 steps over the two buffer sets) must leave exactly what sequential runs leave."""
    from nafae_b200 import _C
    from nafae_b200.pipeline import capture_pipelined_exact
    c, b0, s0 = _step("cfg2", 1, tensor_cores=tensor_cores)
    _, b1, s1 = _step("cfg2", 2, tensor_cores=tensor_cores)
    steps = [s0, s1]
    for st in steps:
        st.run()
    torch.cuda.synchronize()
    want = [[t.clone() for t in (st.rois, st.pooled, st.D_ind, st.loss, st.grad_word)] for st in steps]
    prev = _C.lib.nafae_set_reserved_sms(16)
    try:
        side = [torch.cuda.Stream() for _ in range(3)]
        g = capture_pipelined_exact(steps, 4, side)
        for st in steps:  # scribble over the outputs so that the replays must recompute them
            st.pooled.zero_(); st.rois.zero_(); st.D_ind.zero_(); st.grad_word.zero_(); st.loss.zero_()
        for i in range(3):
            g.replay()
        torch.cuda.synchronize()
    finally:
        _C.lib.nafae_set_reserved_sms(prev)
    for st, w in zip(steps, want):
        for a, t in zip(w, (st.rois, st.pooled, st.D_ind, st.loss, st.grad_word)):
            assert torch.equal(a, t)
    with pytest.raises(ValueError):
        capture_pipelined_exact(steps, 3, side)


@gpu
def test_cfg5_inference_sweep_sample_picks_and_boxes():
    """cfg5 = 10k eval segments sharded over ranks; here a 48-segment shard of rank 1 of 8: picks
    (bit-exact) and the recorded boxes equal the oracle's (postprocess + record_det)."""
    from nafae_b200 import parallel
    from nafae_b200.grounding import ground, postprocess, record_det
    from nafae_b200.model.rpn.proposal_layer import proposal_tail
    c = synth.CONFIGS["cfg1"]
    begin, end = parallel.shard_segments(10000, 1, 8)
    assert (begin, end) == (1250, 2500)
    dev = torch.device("cuda:0")
    for seg in range(begin, begin + 48):
        rs = np.random.RandomState(50000 + seg)
        lens = [int(rs.randint(1, 8))]
        props, scores = synth.proposals(rs, c["Ns"], 300, c["img_h"], c["img_w"])
        vis = synth.embeddings(rs, c["Ns"] * c["Nb"], c["D"])
        word = synth.embeddings(rs, c["Ne"], c["D"])
        rois, _ = proposal_tail(torch.from_numpy(props).to(dev), torch.from_numpy(scores).to(dev),
                                6000, c["Nb"], 0.7)
        D_ind, D_sim, loss = ground(torch.from_numpy(vis).to(dev), torch.from_numpy(word).to(dev),
                                    lens, 1, c["Nb"], c["Ne"], c["Delta"], c["vis_lam"], False)
        D, Ds = postprocess(D_ind, D_sim, 1, c["Ns"], c["Nb"], c["Ne"])
        o_rois, _, _ = ocpu.proposal_tail(props, scores, 6000, c["Nb"], 0.7)
        o_ind, o_sim, _, _ = odvsa.dvsa_forward(torch.from_numpy(vis), torch.from_numpy(word), lens, 1,
                                                c["Nb"], c["Ne"], c["Delta"], c["vis_lam"], "eval")
        oD, oDs = odvsa.postprocess(o_ind.numpy(), o_sim.numpy(), 1, c["Ns"], c["Nb"], c["Ne"])
        n = lens[0]
        np.testing.assert_array_equal(D.cpu().numpy()[:, :, :n], oD[:, :, :n])
        ents = [["e%d" % i for i in range(n)]]
        boxes = rois.view(-1, 5)[:, 1:].cpu().numpy()
        ids = list(range(c["Ns"]))
        got, want = ([], [], [], []), None
        record_det(got[0], got[1], got[2], got[3], c["Nb"], ents, D.cpu().numpy(), Ds.cpu().numpy(), ids, boxes)
        want = odvsa.record_det(c["Nb"], ents, oD, oDs, ids, o_rois.reshape(-1, 5)[:, 1:])
        assert got[0] == want[0] and got[1] == want[1]
        np.testing.assert_array_equal(np.asarray(got[2]), np.asarray(want[2]))


@gpu
def test_symmetric_buffer_alloc_and_world1_allreduce_is_identity():
    """The IPC-mapped gradient bucket: allocation, torch view, world-1 launch (multi-rank behaviour is
    exercised by tests/test_multi_gpu.py and by bench.py --gpus N)."""
    import ctypes
    from nafae_b200 import _C
    from nafae_b200.parallel import _RawCudaArray
    n = 4096
    nbytes = int(_C.lib.nafae_ar_buffer_bytes(n, 1))
    own = ctypes.c_void_p()
    handle = (ctypes.c_ubyte * 64)()
    assert _C.lib.nafae_ar_alloc(nbytes, ctypes.byref(own), handle) == 1
    assert any(bytes(handle))
    off = int(_C.lib.nafae_ar_data_offset())
    buf = torch.as_tensor(_RawCudaArray(own.value + off, n), device="cuda:0")
    assert float(buf.abs().sum()) == 0.0  # zero-filled
    buf.copy_(torch.arange(n, dtype=torch.float32, device="cuda:0"))
    ptrs = (ctypes.c_void_p * 1)(own.value)
    assert _C.lib.nafae_allreduce_avg(ptrs, 0, 1, n, 8, 128, 0, _C.stream()) == 1
    torch.cuda.synchronize()
    assert torch.equal(buf.cpu(), torch.arange(n, dtype=torch.float32))
    del buf
    assert _C.lib.nafae_ar_free(own) == 1


@gpu
@pytest.mark.timeout(100, method="thread")
def test_residency_gate_releases_a_concurrent_branch():
    """nafae_gate_wait: a branch on another stream is released by every gated RoIAlign launch (slab
    kernel and generic-kernel fallback alike), repeatedly, without changing the pooled features.
    (The wait-enqueued-before-the-opener order is what the pipelined step graphs of bench.py use;
    on eager streams a spinning kernel must never precede the FIRST use of another kernel -- lazy
    module loading may synchronise the device -- so here the opener is always enqueued first.)"""
    from nafae_b200 import _C
    from nafae_b200.pipeline import GroundingStep
    dev = torch.device("cuda:0")
    for cfg_name, shape in (("slab", (4, 64, 38, 50)), ("generic_fallback", (2, 6, 37, 50))):
        F, C, H, W = shape
        c = dict(synth.CONFIGS["cfg2"], Na=1, Ns=F, C=C, H=H, W=W, n=300, img_h=H * 16, img_w=W * 16)
        st = GroundingStep(c["Na"], c["Ns"], c["Nb"], c["Ne"], c["D"], C, H, W, c["n"], device=dev)
        st.load(synth.make_batch(c, 11))
        st.run_tail()
        st.run_align(gated=False)
        torch.cuda.synchronize()
        want = st.pooled.clone()
        side = torch.cuda.Stream(dev)
        marks = torch.zeros(4, device=dev)
        for it in range(4):  # several rounds: the gate re-arms itself
            st.pooled.zero_()
            st.run_align(gated=True)
            with torch.cuda.stream(side):
                st.wait_gate(0)
                marks[it:it + 1].fill_(1.0)  # device-side fill: nothing here may block the host
            torch.cuda.synchronize()
            assert float(marks[it]) == 1.0, cfg_name
            assert torch.equal(st.pooled, want), cfg_name
        assert int(st.gate[1]) == 4 and int(st.gate[2]) == 4 and int(st.gate[0]) == 0, cfg_name
    # argument checks: slot range, NULL gate, undersized workspace
    assert _C.lib.nafae_gate_wait(None, 0, _C.stream()) == 0
    assert _C.lib.nafae_gate_wait(_C.ptr(st.gate), 9, _C.stream()) == 0
    assert _C.lib.nafae_roi_align_forward(_C.ptr(st.features), st.scale, st.F, st.R, st.H, st.W, st.C, 7, 7,
                                          _C.POOL_AVG, _C.ptr(st.rois), _C.ptr(st.pooled), 0,
                                          _C.ptr(st.gate), 8, _C.stream()) == 0


@gpu
@pytest.mark.timeout(100, method="thread")
def test_gate_wait_after_unpaired_launches_blocks_until_the_next_open():
    """a gated launch without a waiter (warm-up) must not satisfy a LATER wait.  After
    nafae_gate_sync a waiter really blocks until the concurrent launch opens the gate; without it
    the stale epoch lets it through at once (the round-1 lag, kept visible here)."""
    import time
    from nafae_b200.pipeline import GroundingStep
    dev = torch.device("cuda:0")
    F, C, H, W = 4, 64, 38, 50
    c = dict(synth.CONFIGS["cfg2"], Na=1, Ns=F, C=C, H=H, W=W, n=300, img_h=H * 16, img_w=W * 16)
    st = GroundingStep(c["Na"], c["Ns"], c["Nb"], c["Ne"], c["D"], C, H, W, c["n"], device=dev)
    st.load(synth.make_batch(c, 3))
    st.run_tail()
    side = torch.cuda.Stream(dev)
    marks = torch.zeros(3, device=dev)
    st.run_align(gated=True)           # epoch 1, nobody waits
    with torch.cuda.stream(side):      # (also loads the wait / fill kernels before anything spins)
        side.wait_stream(torch.cuda.current_stream())
        st.wait_gate(0)
        marks[0:1].fill_(1.0)
    torch.cuda.synchronize()
    st.run_align(gated=True)           # epoch 2: UNPAIRED -- slot 0 has only seen epoch 1
    torch.cuda.synchronize()
    with torch.cuda.stream(side):      # stale: passes without any concurrent launch
        st.wait_gate(0)
        marks[1:2].fill_(1.0)
    torch.cuda.synchronize()
    assert float(marks[1]) == 1.0
    st.run_align(gated=True)           # unpaired again, then resynchronise
    st.sync_gate()
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        st.wait_gate(0)
        marks[2:3].fill_(1.0)
    time.sleep(0.2)
    assert not side.query(), "the waiter passed although no launch has opened the gate since the sync"
    want = st.pooled.clone()
    st.run_align(gated=True)           # the launch this wait belongs to
    torch.cuda.synchronize()
    assert float(marks[2]) == 1.0 and torch.equal(st.pooled, want)


@gpu
def test_l1_loss_sign_comes_from_the_forward_workspace():
    """GroundingStep(l1_loss=True) passes no upstream gradient: the backward uses sign(margin_loss)
    (model.py:771) -- same gradients as an explicit +1 for a positive loss."""
    from nafae_b200.pipeline import GroundingStep
    c = synth.CONFIGS["cfg2"]
    b = synth.make_batch("cfg2", 9)
    outs = []
    for l1 in (True, False):
        st = GroundingStep(c["Na"], c["Ns"], c["Nb"], c["Ne"], c["D"], 8, c["H"], c["W"], 64,
                           Delta=c["Delta"], vis_lam=c["vis_lam"], train=True, device="cuda:0", l1_loss=l1)
        st.vis_feats.copy_(torch.from_numpy(b["vis_feats"]))
        st.word_feats.copy_(torch.from_numpy(b["word_feats"]))
        st.lens.copy_(torch.tensor(b["lens"], dtype=torch.int32))
        st.run_head()
        torch.cuda.synchronize()
        assert float(st.loss) > 0
        outs.append((st.grad_word.clone(), st.grad_vis.clone()))
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.allclose(outs[0][1], outs[1][1], rtol=1e-5, atol=1e-7)  # clustering atomics: order


def _raw_stream(s):
    import ctypes
    return ctypes.c_void_p(s.cuda_stream)


@gpu
@pytest.mark.timeout(200, method="thread")
def test_head_and_roi_align_make_progress_on_two_free_sms():
    """Forward progress is not left to convention: with all but TWO SMs held by resident CTAs, the
    kernels whose CTAs wait on other CTAs of the same launch (ground_fwd's phase chain, ground_bwd's
    clustering CTAs behind block 0, the RoIAlign kernel's producer / consumer rings and its dynamic
    claiming) still finish, repeatedly, with the same results."""
    import time
    from nafae_b200 import _C
    c, b, st = _step("cfg2", 21)
    st.run()
    torch.cuda.synchronize()
    want = [t.clone() for t in (st.pooled, st.D_ind, st.D_sim, st.loss, st.grad_word)]
    gv = st.grad_vis.clone()
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    side = torch.cuda.Stream()
    # warm the occupy kernel (module load) before anything spins
    assert _C.lib.nafae_debug_occupy_sms(1, 200 * 1024, 1000, _C.stream()) == 1
    torch.cuda.synchronize()
    for rounds in range(2):
        assert _C.lib.nafae_debug_occupy_sms(sms - 2, 200 * 1024, 400_000_000, _raw_stream(side)) == 1
        time.sleep(0.02)  # the occupying CTAs are resident now
        for _ in range(40):
            st.pooled.zero_()
            st.run_align()
            st.run_head()
        torch.cuda.synchronize()
        for w, t in zip(want, (st.pooled, st.D_ind, st.D_sim, st.loss, st.grad_word)):
            assert torch.equal(w, t)
        assert torch.allclose(gv, st.grad_vis, rtol=1e-5, atol=1e-7)
