"""cfg5 building blocks: G independent segments per launch (batched Na = 1 DVSA forward), device-side
postprocess + record_det + box accuracy (nafae_eval_record) against the oracle's sequential
restatements (oracle/dvsa.py, oracle/eval.py) and the host evaluation (nafae_b200/evaluate.py)."""
import numpy as np
import pytest
import torch

from nafae_b200 import synth
from oracle import cpu as ocpu
from oracle import dvsa as odvsa
from oracle import eval as oeval

gpu = pytest.mark.gpu


@gpu
@pytest.mark.parametrize("G,Ns,Nb,Ne,D,lens", [
    (3, 4, 6, 5, 64, [2, 0, 5]),
    (8, 5, 20, 13, 512, [4, 4, 1, 13, 0, 7, 2, 4]),
    (2, 64, 20, 13, 512, [6, 3]),      # a 64-frame evaluation chunk (reference stepRCNN, model.py:436)
])
def test_eval_step_matches_oracle_per_segment(G, Ns, Nb, Ne, D, lens):
    from nafae_b200 import evaluate
    from nafae_b200.sweep import EvalStep, accuracy_from_counts, dets_from_records
    dev = torch.device("cuda:0")
    C, H, W, n, n_cls = 16, 38, 50, 200, 11
    rs = np.random.RandomState(G * 100 + Ns)
    F = G * Ns
    props, scores = synth.proposals(rs, F, n, H * 16, W * 16)
    feat = synth.conv5_maps(rs, F, C, H, W)
    vis = synth.embeddings(rs, F * Nb, D)
    word = synth.embeddings(rs, G * Ne, D)
    cls = np.full((G, Ne), -1, np.int32)
    for g in range(G):
        cls[g, :lens[g]] = rs.choice(n_cls, lens[g], replace=False) if lens[g] <= n_cls else rs.randint(0, n_cls, lens[g])
    es = EvalStep(G, Ns, Nb, Ne, D, C, H, W, n, n_cls, device=dev)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    out = dict(image_ids=torch.empty((G, Ns, Ne), dtype=torch.int64, device=dev),
               box_rows=torch.empty((G, Ns, Ne), dtype=torch.int64, device=dev),
               boxes=torch.empty((G, Ns, Ne, 4), device=dev), confs=torch.empty((G, Ns, Ne), device=dev))
    # ground truth: the box the oracle picks, jittered (some hits, some misses)
    o_rois, _, _ = ocpu.proposal_tail(props, scores, 6000, Nb, 0.7)
    rows = o_rois.reshape(-1, 5)[:, 1:]
    gt = np.zeros((G, Ns, Ne, 4))
    want_D, want_sim, want_loss = [], [], []
    for g in range(G):
        v = torch.from_numpy(vis[g * Ns * Nb:(g + 1) * Ns * Nb])
        w = torch.from_numpy(word[g * Ne:(g + 1) * Ne])
        o_ind, o_sim, o_loss, _ = odvsa.dvsa_forward(v, w, [lens[g]], 1, Nb, Ne, 5.0, 1.0, "eval")
        oD, oS = odvsa.postprocess(o_ind.numpy(), o_sim.numpy(), 1, Ns, Nb, Ne)
        want_D.append(oD[0] + g * Ns * Nb)
        want_sim.append(oS[0])
        want_loss.append(float(o_loss))
        gt[g] = rows[want_D[-1]] + rs.uniform(-25, 25, (Ns, Ne, 4))
    gt[..., 2:] = np.maximum(gt[..., 2:], gt[..., :2] + 1)
    counts = torch.zeros((2, n_cls), dtype=torch.int32, device=dev)
    for rep in range(2):  # twice: workspaces / counters of the kernels must be left reusable
        counts.zero_()
        es.run(t(feat), t(props), t(scores), t(vis), t(word), t(np.asarray(lens, np.int32)), 1000, out=out,
               gt_boxes=t(gt), gt_classes=t(cls), class_match=counts[0], class_count=counts[1])
        torch.cuda.synchronize()
        np.testing.assert_array_equal(es.rois.cpu().numpy(), o_rois)
        ids = out["image_ids"].cpu().numpy()
        for g in range(G):
            k = lens[g]
            np.testing.assert_array_equal(out["box_rows"][g].cpu().numpy()[:, :k], want_D[g][:, :k])
            np.testing.assert_array_equal(out["boxes"][g].cpu().numpy()[:, :k], rows[want_D[g][:, :k]])
            np.testing.assert_allclose(out["confs"][g].cpu().numpy()[:, :k], want_sim[g][:, :k], rtol=1e-4, atol=1e-5)
            assert (ids[g][:, :k] == 1000 + g * Ns + np.arange(Ns)[:, None]).all() and (ids[g][:, k:] == -1).all()
            np.testing.assert_allclose(float(es.loss[g]), want_loss[g], rtol=1e-4)
    # accuracy counters == host evaluation == sequential oracle on the recorded detections
    classes = ["c%d" % i for i in range(n_cls)]
    local = out["image_ids"].clone()
    local[local >= 0] -= 1000
    dets = dets_from_records(local, out["boxes"], out["confs"], lambda g, e: classes[int(cls[g, e])])
    recs = [dict(label=[classes[int(cls[g, e])] for e in range(lens[g])],
                 bbox=[gt[g, f, e] for e in range(lens[g])], thr=[0.5] * lens[g])
            for g in range(G) for f in range(Ns)]
    # duplicate labels (possible only in the lens > n_cls case) take the reference's sequential path,
    # which the aligned device reduction does not model
    if all(len(set(cls[g, :lens[g]])) == lens[g] for g in range(G)) and sum(lens) > 0:
        host = evaluate.box_accuracy_details(recs, dets, classes)
        orc = oeval.box_accuracy(recs, dets, classes)
        cnt = counts.cpu().numpy()
        np.testing.assert_array_equal(cnt[0], host["class_match_count"])
        np.testing.assert_array_equal(cnt[1], host["class_count"])
        np.testing.assert_array_equal(cnt[0], orc["class_match_count"])
        np.testing.assert_array_equal(cnt[0], oeval.phrase_accuracy(recs, dets, classes)["class_match_count"])
        macro, micro = accuracy_from_counts(counts[0], counts[1])
        assert abs(macro - host["macro"]) < 1e-12 and abs(micro - host["micro"]) < 1e-12
        assert 0 < cnt[0].sum() < cnt[1].sum()  # the jitter produced both hits and misses


@gpu
def test_batched_groups_equal_separate_na1_calls():
    """nafae_ground_forward_batched(groups = G, Na = 1) == G calls of nafae_ground_forward."""
    from nafae_b200.grounding import ground
    from nafae_b200 import _C
    dev = torch.device("cuda:0")
    G, Ns, Nb, Ne, D = 5, 7, 20, 13, 512
    rs = np.random.RandomState(3)
    vis = torch.from_numpy(synth.embeddings(rs, G * Ns * Nb, D)).to(dev)
    word = torch.from_numpy(synth.embeddings(rs, G * Ne, D)).to(dev)
    lens = [3, 13, 0, 1, 6]
    lt = torch.tensor(lens, dtype=torch.int32, device=dev)
    D_ind = torch.empty((G, Ns, Ne), dtype=torch.int64, device=dev)
    D_sim = torch.empty((G, Ns, Ne), device=dev)
    loss = torch.empty((G,), device=dev)
    wsb = int(_C.lib.nafae_ground_workspace_bytes(1, Ns, Nb, Ne, D))
    ws = torch.zeros((wsb * G // 4,), dtype=torch.int32, device=dev)
    st = _C.lib.nafae_ground_forward_batched(_C.ptr(vis), _C.ptr(word), _C.ptr(lt), G, 1, Ns, Nb, Ne, D, 5.0,
                                             1.0, 0, _C.ptr(D_ind), _C.ptr(D_sim), _C.ptr(loss), _C.ptr(ws),
                                             ws.numel() * 4, _C.stream())
    assert st == 1, _C.last_error()
    for g in range(G):
        i, s, l = ground(vis[g * Ns * Nb:(g + 1) * Ns * Nb], word[g * Ne:(g + 1) * Ne], [lens[g]], 1, Nb, Ne,
                         5.0, 1.0, False)
        assert torch.equal(D_ind[g], i) and torch.equal(D_sim[g], s) and torch.equal(loss[g], l)
    # too small a workspace for the groups is an argument error
    assert _C.lib.nafae_ground_forward_batched(_C.ptr(vis), _C.ptr(word), _C.ptr(lt), G, 1, Ns, Nb, Ne, D, 5.0,
                                               1.0, 0, _C.ptr(D_ind), _C.ptr(D_sim), _C.ptr(loss), _C.ptr(ws),
                                               wsb, _C.stream()) == 0
