"""TEST INFRASTRUCTURE ONLY (parity oracle; never imported by the product path).

Sequential CPU restatement of the reference's grounding evaluation,
lib/datasets/youcook_eval.py: `phrase_accuracy` (:135-237, "query accuracy"), `box_accuracy`
(:241-336) and `evaluate_box` (:408-413).  Pinned by tests/golden/eval_*.npz, which hold the outputs
of the reference's own function bodies executed on seeded inputs (tests/golden/make_eval_golden.py).

Quirks of the reference that are kept because they change results:
  * both functions permute the confidences TWICE (`obj_confs[order]` at :153 and again at :157 /
    :259 and :263), so within an image the detections are visited in the order of somebody else's
    confidence whenever `order` is not the identity;
  * `phrase_accuracy` looks the class index up only when a label is seen for the first time in an
    image (:198-201); a later detection with an already-seen, still unmatched label books its match
    under whatever class index was looked up last (:224);
  * only images `0 .. max(det img_ids)` are visited (:158, :264): ground truth of later images is
    not counted;
  * macro accuracy divides by `count + 1e-6` and averages over ALL classes, present or not (:228-229, :327-328).
"""
import numpy as np


def _overlap(det_box, gt_box):
    """IoU with the +1 pixel convention, or None when the boxes do not intersect (:206-219).
    Same expressions (and therefore the same NumPy scalar promotions) as the reference."""
    left = np.max((det_box[0], gt_box[0]))
    top = np.max((det_box[1], gt_box[1]))
    right = np.min((det_box[2], gt_box[2]))
    bottom = np.min((det_box[3], gt_box[3]))
    iw = right - left + 1
    ih = bottom - top + 1
    if not (iw > 0 and ih > 0):
        return None
    union = (det_box[2] - det_box[0] + 1.) * (det_box[3] - det_box[1] + 1.) + \
        (gt_box[2] - gt_box[0] + 1.) * (gt_box[3] - gt_box[1] + 1.) - iw * ih
    return iw * ih / union


def _per_image_detections(dets):
    """:143-177 / :249-283: sort by image id, (double-permuted) confidences, one list of
    (label, box) per image in descending order of that confidence."""
    img_ids = np.array(dets[0])
    labels = np.array(dets[1])
    boxes = np.array(dets[2])
    confs = np.array(dets[3])
    order = np.argsort(img_ids)
    img_ids, labels, boxes = img_ids[order], labels[order], boxes[order]
    confs = confs[order][order]
    n_imgs = int(np.max(img_ids)) + 1
    cells = [None] * n_imgs
    begin = 0
    for k in range(len(img_ids)):
        if k == len(img_ids) - 1 or img_ids[k + 1] != img_ids[begin]:
            rank = np.argsort(-confs[begin:k + 1])
            cells[int(img_ids[begin])] = [(labels[begin + r], boxes[begin + r]) for r in rank]
            begin = k + 1
    return cells


def _summary(match, count):
    per_class = match / (count + 1e-6)
    return dict(macro=float(np.mean(per_class)), micro=float(np.sum(match) / np.sum(count)),
                class_match_count=match, class_count=count)


def phrase_accuracy(recs, dets, class_list):
    """youcook_eval.py:135-237.  One trial per (image, grounded label that is annotated in that
    image); a hit if any detection / ground-truth pair of that label overlaps by >= its threshold."""
    cells = _per_image_detections(dets)
    match = np.zeros(len(class_list), dtype=int)
    count = np.zeros(len(class_list), dtype=int)
    class_ind = None
    for img_id, cell in enumerate(cells):
        if cell is None:
            continue
        rec = recs[img_id]
        state = {}  # label -> matched?
        for label, box in cell:
            for gt_label, gt_box, thr in zip(rec['label'], rec['bbox'], rec['thr']):
                if label != gt_label:
                    continue
                if label not in state:
                    state[label] = False
                    class_ind = class_list.index(gt_label)
                    count[class_ind] += 1
                elif state[label]:
                    continue
                ov = _overlap(box, gt_box)
                if ov is not None and ov >= thr:
                    match[class_ind] += 1  # the index looked up LAST (quirk, see module docstring)
                    state[label] = True
    return _summary(match, count)


def box_accuracy(recs, dets, class_list):
    """youcook_eval.py:241-336.  One trial per annotated box; a hit if any detection of the same
    label in that image overlaps it by >= its threshold."""
    cells = _per_image_detections(dets)
    match = np.zeros(len(class_list), dtype=int)
    count = np.zeros(len(class_list), dtype=int)
    for img_id, cell in enumerate(cells):
        rec = recs[img_id]
        for gt_label, gt_box, thr in zip(rec['label'], rec['bbox'], rec['thr']):
            class_ind = class_list.index(gt_label)
            count[class_ind] += 1
            if cell is None:
                continue
            for label, box in cell:
                if label != gt_label:
                    continue
                ov = _overlap(box, gt_box)
                if ov is not None and ov >= thr:
                    match[class_ind] += 1
                    break
    return _summary(match, count)


def evaluate_box(recs, dets, class_list):
    """youcook_eval.py:408-413: query-level then box-level; returns the macro BOX accuracy."""
    phrase_accuracy(recs, dets, class_list)
    return box_accuracy(recs, dets, class_list)['macro']
