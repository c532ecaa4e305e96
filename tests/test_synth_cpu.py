"""Synthetic workload generator: shapes, determinism, NMS input contract."""
import numpy as np

from nafae_b200 import synth


def test_batches_are_deterministic_and_shaped():
    a = synth.make_batch("cfg2_real", 5)
    b = synth.make_batch("cfg2_real", 5)
    for k in a:
        if k != "lens":
            assert np.array_equal(a[k], b[k])
    c = synth.CONFIGS["cfg2_real"]
    F = c["Na"] * c["Ns"]
    assert a["features"].shape == (F, c["C"], c["H"], c["W"]) and a["features"].min() >= 0
    assert a["proposals"].shape == (F, c["n"], 4) and a["scores"].shape == (F, c["n"])
    assert a["vis_feats"].shape == (F * c["Nb"], c["D"]) and np.abs(a["vis_feats"]).max() <= 1
    assert len(a["lens"]) == c["Na"] and sum(a["lens"]) > 0 and max(a["lens"]) <= c["Ne"]


def test_proposals_respect_the_nms_contract():
    rs = np.random.RandomState(0)
    p, s = synth.proposals(rs, 3, 500, 608, 800)
    assert (np.diff(s, axis=1) <= 0).all()            # score-descending
    assert (p[..., 0] >= 0).all() and (p[..., 2] <= 799).all()
    assert (p[..., 1] >= 0).all() and (p[..., 3] <= 607).all()
    assert (p[..., 2] >= p[..., 0]).all() and (p[..., 3] >= p[..., 1]).all()
