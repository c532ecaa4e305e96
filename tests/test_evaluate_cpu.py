"""Evaluation parity (host-side component, no GPU needed): oracle/eval.py and nafae_b200/evaluate.py
against fixtures produced by the reference's own youcook_eval.py functions
(tests/golden/make_eval_golden.py), and the vectorised implementation against the oracle on larger
seeded cases (integer counts: bit-exact)."""
import contextlib
import glob
import io
import os
import sys

import numpy as np
import pytest

from _golden import GOLDEN
sys.path.insert(0, GOLDEN)
from make_eval_golden import make_case, to_reference_inputs  # noqa: E402
from nafae_b200 import evaluate  # noqa: E402
from oracle import eval as oeval  # noqa: E402

FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "eval_*.npz")))


def _id(p):
    return os.path.basename(p)[len("eval_"):-4]


def test_fixtures_exist():
    assert len(FIXTURES) >= 6


@pytest.mark.parametrize("path", FIXTURES, ids=_id)
def test_oracle_reproduces_the_reference_outputs(path):
    z = np.load(path)
    recs, dets, class_list = to_reference_inputs(z)
    p = oeval.phrase_accuracy(recs, dets, class_list)
    b = oeval.box_accuracy(recs, dets, class_list)
    assert p["macro"] == float(z["phrase_macro"]) and b["macro"] == float(z["box_macro"])
    assert 'micro query accuracy: {:0.2%}'.format(p["micro"]) == str(z["phrase_printed"][1])
    assert 'micro box accuracy: {:0.2%}'.format(b["micro"]) == str(z["box_printed"][1])


@pytest.mark.parametrize("path", FIXTURES, ids=_id)
def test_vectorised_evaluation_reproduces_the_reference_outputs(path):
    z = np.load(path)
    recs, dets, class_list = to_reference_inputs(z)
    for fn, key in ((evaluate.phrase_accuracy, "phrase"), (evaluate.box_accuracy, "box")):
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            macro = fn(recs, dets, class_list)
        assert macro == float(z[key + "_macro"])                       # same float, not "close"
        assert buf.getvalue().strip().splitlines() == [str(s) for s in z[key + "_printed"]]
    with contextlib.redirect_stdout(io.StringIO()):
        assert evaluate.evaluate_box(recs, dets, class_list) == float(z["box_macro"])


@pytest.mark.parametrize("kw", [
    dict(seed=11, n_imgs=400, n_cls=67),
    dict(seed=12, n_imgs=400, n_cls=9, dup_labels=True),
    dict(seed=13, n_imgs=300, n_cls=7, dup_labels=True, shuffle=True),
    dict(seed=14, n_imgs=300, n_cls=20, shuffle=True, ties=True),
    dict(seed=15, n_imgs=2, n_cls=2, dup_labels=True),
], ids=lambda kw: "seed%d" % kw["seed"])
def test_vectorised_counts_equal_the_sequential_oracle(kw):
    z = make_case(**kw)
    recs, dets, class_list = to_reference_inputs(z)
    for got, want in ((evaluate.phrase_accuracy_details(recs, dets, class_list),
                       oeval.phrase_accuracy(recs, dets, class_list)),
                      (evaluate.box_accuracy_details(recs, dets, class_list),
                       oeval.box_accuracy(recs, dets, class_list))):
        np.testing.assert_array_equal(got["class_count"], want["class_count"])
        np.testing.assert_array_equal(got["class_match_count"], want["class_match_count"])
        assert got["macro"] == want["macro"] and got["micro"] == want["micro"]


def test_edge_cases_follow_the_reference():
    class_list = ["a", "b"]
    recs = [dict(label=["a"], bbox=[[0, 0, 9, 9]], thr=[0.5], img_ids=[0]),
            dict(label=["b"], bbox=[[0, 0, 9, 9]], thr=[0.5], img_ids=[1])]
    box = np.array([0, 0, 9, 9], np.float32)
    # ground truth after the last image with a detection is not counted (youcook_eval.py:264)
    dets = [[0], ["a"], [box], [np.float64(0.3)]]
    r = evaluate.box_accuracy_details(recs, dets, class_list)
    assert r["class_count"].tolist() == [1, 0] and r["class_match_count"].tolist() == [1, 0]
    # a grounded label outside class_list never matches and is not a trial
    dets = [[0, 1], ["zzz", "b"], [box, box], [np.float64(0.3), np.float64(0.1)]]
    r = evaluate.phrase_accuracy_details(recs, dets, class_list)
    assert r["class_count"].tolist() == [0, 1] and r["class_match_count"].tolist() == [0, 1]
    # touching but not overlapping boxes: iw == 0 -> no match; overlap exactly at the threshold matches
    far = np.array([10, 0, 19, 9], np.float32)
    half = np.array([0, 0, 9, 4], np.float32)
    for b, want in ((far, 0), (half, 1)):
        r = evaluate.box_accuracy_details(recs, [[0], ["a"], [b], [np.float64(1.0)]], class_list)
        assert r["class_match_count"].tolist() == [want, 0]
    with pytest.raises(ValueError):  # gt label missing from class_list: class_list.index raises
        evaluate.box_accuracy_details([dict(label=["q"], bbox=[[0, 0, 1, 1]], thr=[0.5], img_ids=[0])],
                                      [[0], ["a"], [box], [np.float64(0.0)]], class_list)
    with pytest.raises(ValueError):  # no detections at all
        evaluate.box_accuracy_details(recs, [[], [], [], []], class_list)


def test_result_file_round_trip(tmp_path):
    z = np.load(FIXTURES[0])
    _, dets, _ = to_reference_inputs(z)
    path = str(tmp_path / "ground_res_val_1_1_0.pkl")
    evaluate.save_dets(path, dets)
    back = evaluate.load_dets(path)
    assert back[0] == dets[0] and back[1] == dets[1]
    np.testing.assert_array_equal(np.array(back[2]), np.array(dets[2]))
    np.testing.assert_array_equal(np.array(back[3]), np.array(dets[3]))
    import pickle
    with open(path, "rb") as f:
        raw = pickle.load(f)
    assert isinstance(raw, list) and len(raw) == 4        # model.py:972: [img_inds, labels, bboxes, confs]
    with open(path, "wb") as f:
        pickle.dump([[1], [2]], f)
    with pytest.raises(ValueError):
        evaluate.load_dets(path)


def test_record_det_tensors_equals_the_reference_loop():
    """Vectorised record_det (grounding.record_det_tensors) against the reference's triple loop
    (oracle.dvsa.record_det, model.py:477-487) -- CPU tensors here, the same code runs on the device."""
    import torch
    from nafae_b200.grounding import record_det_tensors
    from oracle import dvsa as odvsa
    rs = np.random.RandomState(3)
    Na, Ns, Nb, Ne = 4, 5, 20, 13
    lens = [3, 0, 13, 1]
    D = np.zeros((Na, Ns, Ne), dtype=np.int64)
    for a in range(Na):
        for s in range(Ns):
            D[a, s] = rs.randint(0, Nb, size=Ne) + a * Ns * Nb + s * Nb   # postprocess, model.py:470
    D_sim = rs.randn(Na, Ns, Ne)
    img_ids = list(range(100, 100 + Na * Ns))
    boxes = rs.uniform(0, 200, size=(Na * Ns * Nb, 4)).astype(np.float32)
    ents = [["e%d_%d" % (a, e) for e in range(n)] for a, n in enumerate(lens)]
    want = odvsa.record_det(Nb, ents, D, D_sim, img_ids, boxes)
    img, seg, ent, bx, cf = record_det_tensors(torch.from_numpy(D), torch.from_numpy(D_sim), lens,
                                               torch.tensor(img_ids), torch.from_numpy(boxes), Nb)
    assert img.tolist() == list(want[0])
    assert [ents[a][e] for a, e in zip(seg.tolist(), ent.tolist())] == list(want[1])
    np.testing.assert_array_equal(bx.numpy(), np.asarray(want[2]))
    np.testing.assert_array_equal(cf.numpy(), np.asarray(want[3]))
