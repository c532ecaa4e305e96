"""Seeded input builders shared by the golden-fixture generator and the parity tests."""
import numpy as np

from nafae_b200 import synth


def nms_cases():
    """name -> (dets (n,5) f32 sorted by score desc, thresh)."""
    cases = {}
    rs = np.random.RandomState(21)
    for name, n, thr in (("n300", 300, 0.7), ("n2352", 2352, 0.7), ("n64", 64, 0.5),
                         ("n65", 65, 0.7), ("n1", 1, 0.7), ("n129_t03", 129, 0.3)):
        p, s = synth.proposals(rs, 1, n, 608, 800)
        cases[name] = (np.concatenate([p[0], s[0][:, None]], 1).astype(np.float32), thr)
    # exact duplicates, nested and touching boxes, degenerate (x2 < x1) boxes, ties at IoU==thresh
    d = np.array([[10, 10, 50, 50], [10, 10, 50, 50], [10, 10, 50, 49], [51, 10, 90, 50],
                  [0, 0, 0, 0], [0, 0, 0, 0], [20, 20, 10, 10], [20, 20, 10, 10],
                  [0, 0, 9, 9], [0, 0, 9, 19], [5, 5, 5, 5], [0, 0, 799, 607]], np.float32)
    sc = np.linspace(1, 0.1, len(d)).astype(np.float32)
    cases["edge"] = (np.concatenate([d, sc[:, None]], 1), 0.5)
    # dense small integer grid: many IoUs equal simple fractions, some exactly at the threshold
    g = []
    for x in range(0, 12, 2):
        for y in range(0, 12, 3):
            g.append([x, y, x + 9, y + 9])
    g = np.array(g, np.float32)
    sc = np.linspace(1, 0.1, len(g)).astype(np.float32)
    cases["grid_t05"] = (np.concatenate([g, sc[:, None]], 1), 0.5)
    return cases


def roi_cases():
    """name -> dict(features (B,C,H,W), rois (R,5), scale, ah, aw)."""
    cases = {}
    rs = np.random.RandomState(31)
    B, C, H, W = 2, 6, 9, 11
    feat = rs.standard_normal((B, C, H, W)).astype(np.float32)
    s = 1.0 / 16.0
    iw, ih = W * 16, H * 16
    rois = np.array([
        [0, 0, 0, 0, 0],                      # zero padded proposal (proposal_layer.py:127)
        [1, 0, 0, 0, 0],
        [0, 10.5, 20.25, 100.75, 90.5],
        [1, 0, 0, iw - 1, ih - 1],            # whole image: last samples fall in [H-1, H)
        [0, 30, 40, 30, 40],                  # single pixel
        [1, 100, 60, 40, 20],                 # malformed: x2 < x1
        [0, -40, -30, 60, 50],                # partly outside (negative start)
        [1, 120, 100, 400, 300],              # runs past the map: samples >= H / W -> 0
        [0, iw - 17, ih - 17, iw - 1, ih - 1],
        [1, 3.3, 7.7, 150.2, 130.9],
    ], np.float32)
    cases["small_8x8"] = dict(features=feat, rois=rois, scale=s, ah=8, aw=8)
    cases["small_3x5"] = dict(features=feat, rois=rois, scale=s, ah=3, aw=5)
    cases["small_7x7_scale8"] = dict(features=feat, rois=rois * np.array([1, .5, .5, .5, .5],
                                                                          np.float32),
                                     scale=1.0 / 8.0, ah=7, aw=7)
    # map of the benchmark shape (38x50), few channels
    feat2 = np.maximum(rs.standard_normal((1, 8, 38, 50)), 0).astype(np.float32)
    p, _ = synth.proposals(rs, 1, 12, 608, 800)
    rois2 = np.concatenate([np.zeros((12, 1), np.float32), p[0]], 1)
    cases["map38x50_8x8"] = dict(features=feat2, rois=rois2, scale=s, ah=8, aw=8)
    # reference-real 14x14 map
    feat3 = np.maximum(rs.standard_normal((3, 5, 14, 14)), 0).astype(np.float32)
    p, _ = synth.proposals(rs, 3, 5, 224, 224)
    rois3 = np.concatenate([np.repeat(np.arange(3, dtype=np.float32), 5)[:, None],
                            p.reshape(-1, 4)], 1)
    rois3 = rois3[rs.permutation(len(rois3))]  # unsorted batch indices
    cases["map14x14_8x8"] = dict(features=feat3, rois=rois3, scale=s, ah=8, aw=8)
    return cases


def top_diff_for(case, out_h, out_w, seed=41):
    rs = np.random.RandomState(seed)
    R = case["rois"].shape[0]
    C = case["features"].shape[1]
    return rs.standard_normal((R, C, out_h, out_w)).astype(np.float32)
