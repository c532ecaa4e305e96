"""Grounding evaluation: box accuracy and query ("phrase") accuracy, vectorised.

Mirrors the reference's evaluation surface -- lib/datasets/youcook_eval.py: `phrase_accuracy`
(:135-237), `box_accuracy` (:241-336), `evaluate_box` (:408-413) -- with the same arguments
(`recs`: one dict per image with 'label' / 'bbox' / 'thr' lists as `parse_gt` builds them, :78-110;
`dets`: the four parallel lists `[img_ids, labels, bboxes, confs]` that `record_det` fills,
model.py:477-487, 972), the same printed lines and the same return value (macro accuracy), plus the
result-file format of model.py:972-983 (`save_dets` / `load_dets`).

The reference walks Python loops over images x detections x ground-truth boxes (minutes for the
cfg5 sweep of 10 000 segments); here every (detection, ground truth) pair of the same image and
class is materialised once with a sort-join and scored with array arithmetic.  This is host-side
NumPy like the reference's own evaluation: it is bookkeeping after the device has produced the
picks, not a fallback for a device kernel.

Results are identical to the reference's, including two of its quirks (see oracle/eval.py):
ground truth of images after the last image that has a detection is not counted, and an image in
which one label is grounded twice goes through the reference's sequential bookkeeping (visit order
from the doubly permuted confidences, match booked under the class index looked up last).
"""
import pickle

import numpy as np


def _class_ids(class_list):
    ids = {}
    for i, name in enumerate(class_list):
        ids.setdefault(name, i)  # list.index semantics: first occurrence
    return ids


def _arrays(recs, dets, class_list):
    ids = _class_ids(class_list)
    d_img = np.array(dets[0])
    if d_img.size == 0:
        raise ValueError("no detections (the reference fails on np.max of an empty array, too)")
    d_cls = np.array([ids.get(label, -1) for label in dets[1]], dtype=np.int64)
    d_box = np.array(dets[2])
    d_conf = np.array(dets[3])
    n_imgs = int(np.max(d_img)) + 1  # youcook_eval.py:158, :264
    g_img, g_cls, g_box, g_thr = [], [], [], []
    for img_id in range(n_imgs):
        rec = recs[img_id]
        for label, box, thr in zip(rec['label'], rec['bbox'], rec['thr']):
            if label not in ids:
                raise ValueError("%r is not in list" % (label,))  # class_list.index(gt_label)
            g_img.append(img_id)
            g_cls.append(ids[label])
            g_box.append(box)
            g_thr.append(thr)
    g_box = np.array(g_box).reshape(-1, 4)
    return (d_img.astype(np.int64), d_cls, d_box.reshape(-1, 4), d_conf,
            np.array(g_img, dtype=np.int64), np.array(g_cls, dtype=np.int64), g_box,
            np.array(g_thr, dtype=np.float64), n_imgs)


def _pairs(d_img, d_cls, g_img, g_cls, n_cls):
    """All (detection index, ground-truth index) pairs with equal image and class (sort-join).
    Ground truth keeps its per-image order inside a key (stable sort)."""
    if g_img.size == 0:
        z = np.zeros(0, dtype=np.int64)
        return z, z
    g_key = g_img * n_cls + g_cls
    g_order = np.argsort(g_key, kind="stable")
    g_sorted = g_key[g_order]
    live = np.flatnonzero(d_cls >= 0)  # a label outside class_list can never equal a gt label
    d_key = d_img[live] * n_cls + d_cls[live]
    lo = np.searchsorted(g_sorted, d_key, side="left")
    hi = np.searchsorted(g_sorted, d_key, side="right")
    cnt = hi - lo
    pd = np.repeat(live, cnt)
    start = np.repeat(lo, cnt)
    within = np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt)
    pg = g_order[start + within]
    return pd, pg


def _hits(d_box, g_box, g_thr):
    """overlap >= thr for aligned rows of boxes; same expressions and dtypes as
    youcook_eval.py:206-221 (det area in the detections' own precision, the rest in float64)."""
    left = np.maximum(d_box[:, 0], g_box[:, 0]).astype(np.float64)
    top = np.maximum(d_box[:, 1], g_box[:, 1]).astype(np.float64)
    right = np.minimum(d_box[:, 2], g_box[:, 2]).astype(np.float64)
    bottom = np.minimum(d_box[:, 3], g_box[:, 3]).astype(np.float64)
    iw = right - left + 1
    ih = bottom - top + 1
    ok = (iw > 0) & (ih > 0)
    union = (d_box[:, 2] - d_box[:, 0] + 1.) * (d_box[:, 3] - d_box[:, 1] + 1.) + \
        (g_box[:, 2] - g_box[:, 0] + 1.) * (g_box[:, 3] - g_box[:, 1] + 1.) - iw * ih
    with np.errstate(divide="ignore", invalid="ignore"):
        ov = iw * ih / union
    return ok & (ov >= g_thr)


def _summary(match, count):
    per_class = match / (count + 1e-6)
    return dict(macro=float(np.mean(per_class)), micro=float(np.sum(match) / np.sum(count)),
                class_match_count=match, class_count=count)


def box_accuracy_details(recs, dets, class_list):
    d_img, d_cls, d_box, _, g_img, g_cls, g_box, g_thr, _ = _arrays(recs, dets, class_list)
    n_cls = len(class_list)
    pd, pg = _pairs(d_img, d_cls, g_img, g_cls, n_cls)
    hit = _hits(d_box[pd], g_box[pg], g_thr[pg])
    matched = np.zeros(g_img.size, dtype=bool)
    matched[pg[hit]] = True
    count = np.bincount(g_cls, minlength=n_cls).astype(int)
    match = np.bincount(g_cls[matched], minlength=n_cls).astype(int)
    return _summary(match, count)


def _visit_order(d_img, d_conf):
    """For every detection its position in the reference's walk: sorted by image, inside an image by
    descending confidence AFTER the double permutation of youcook_eval.py:153+157."""
    order = np.argsort(d_img)
    img_sorted = d_img[order]
    conf2 = d_conf[order][order]
    visit = np.empty(d_img.size, dtype=np.int64)
    bounds = np.flatnonzero(np.r_[True, img_sorted[1:] != img_sorted[:-1], True])
    for b, e in zip(bounds[:-1], bounds[1:]):
        visit[b:e] = order[b + np.argsort(-conf2[b:e])]
    return visit, img_sorted, bounds


def phrase_accuracy_details(recs, dets, class_list):
    d_img, d_cls, d_box, d_conf, g_img, g_cls, g_box, g_thr, _ = _arrays(recs, dets, class_list)
    n_cls = len(class_list)
    pd, pg = _pairs(d_img, d_cls, g_img, g_cls, n_cls)
    hit = _hits(d_box[pd], g_box[pg], g_thr[pg])
    count = np.zeros(n_cls, dtype=int)
    match = np.zeros(n_cls, dtype=int)
    if pd.size == 0:
        return _summary(match, count)
    key = d_img[pd] * n_cls + d_cls[pd]  # one trial per (image, grounded label annotated there)
    # images where one active label is grounded by more than one detection: sequential bookkeeping
    act_d, act_first = np.unique(pd, return_index=True)
    act_key = key[act_first]
    uk, per_key = np.unique(act_key, return_counts=True)
    dup_imgs = np.unique(uk[per_key > 1] // n_cls)
    seq = np.isin(d_img[pd], dup_imgs)
    # --- vectorised part
    k2, inv = np.unique(key[~seq], return_inverse=True)
    any_hit = np.zeros(k2.size, dtype=bool)
    np.logical_or.at(any_hit, inv, hit[~seq])
    cls2 = (k2 % n_cls).astype(np.int64)
    count += np.bincount(cls2, minlength=n_cls).astype(int)
    match += np.bincount(cls2[any_hit], minlength=n_cls).astype(int)
    # --- sequential part (youcook_eval.py:185-226 on the affected images only)
    if dup_imgs.size:
        visit, img_sorted, bounds = _visit_order(d_img, d_conf)
        starts = {int(img_sorted[b]): (b, e) for b, e in zip(bounds[:-1], bounds[1:])}
        g_by_img = {}
        for gi in np.flatnonzero(np.isin(g_img, dup_imgs)):
            g_by_img.setdefault(int(g_img[gi]), []).append(int(gi))
        class_ind = -1
        for img in dup_imgs.tolist():
            b, e = starts[img]
            state = {}
            for di in visit[b:e].tolist():
                c = int(d_cls[di])
                if c < 0:
                    continue
                for gi in g_by_img.get(img, ()):
                    if int(g_cls[gi]) != c:
                        continue
                    if c not in state:
                        state[c] = False
                        class_ind = c
                        count[c] += 1
                    elif state[c]:
                        continue
                    if _hits(d_box[di:di + 1], g_box[gi:gi + 1], g_thr[gi:gi + 1])[0]:
                        match[class_ind] += 1
                        state[c] = True
    return _summary(match, count)


def phrase_accuracy(recs, dets, class_list, verbose=True):
    """youcook_eval.py:135-237; returns the macro query accuracy."""
    r = phrase_accuracy_details(recs, dets, class_list)
    if verbose:
        print('macro query accuracy: {:0.2%}'.format(r['macro']))
        print('micro query accuracy: {:0.2%}'.format(r['micro']))
    return r['macro']


def box_accuracy(recs, dets, class_list, verbose=True):
    """youcook_eval.py:241-336; returns the macro box accuracy."""
    r = box_accuracy_details(recs, dets, class_list)
    if verbose:
        print('macro box accuracy: {:0.2%}'.format(r['macro']))
        print('micro box accuracy: {:0.2%}'.format(r['micro']))
    return r['macro']


def evaluate_box(recs, dets, class_list, verbose=True):
    """youcook_eval.py:408-413: query-level, then box-level; returns the macro box accuracy."""
    phrase_accuracy(recs, dets, class_list, verbose)
    return box_accuracy(recs, dets, class_list, verbose)


# ------------------------------------------------------------------ result files ----
def save_dets(path, dets):
    """The result pickle of model.py:972-983: `[img_inds, obj_labels, obj_bboxes, obj_confs]`, readable
    by the reference's evaluation entry point (youcook_eval.py:420-453)."""
    img_inds, labels, boxes, confs = dets
    with open(path, 'wb') as f:
        pickle.dump([list(img_inds), list(labels), list(boxes), list(confs)], f)


def load_dets(path):
    with open(path, 'rb') as f:
        dets = pickle.load(f)
    if not (isinstance(dets, (list, tuple)) and len(dets) == 4 and
            len({len(x) for x in dets}) == 1):
        raise ValueError("not a grounding result file: expected four parallel lists")
    return list(dets)
