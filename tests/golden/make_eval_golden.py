"""Generates tests/golden/eval_*.npz by EXECUTING the reference's own evaluation functions
(lib/datasets/youcook_eval.py: phrase_accuracy :135-237, box_accuracy :241-336) on seeded synthetic
ground truth / detections.  Run in the build container (needs /root/reference); the fixtures travel.

The module itself cannot be imported (nltk, tqdm at import time), so the two function definitions
are cut out of the source file with `ast` and exec'd with numpy only -- not a single token changed.

    python tests/golden/make_eval_golden.py
"""
import ast
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/lib/datasets/youcook_eval.py"


def reference_functions():
    src = open(REF).read()
    tree = ast.parse(src)
    ns = {"np": np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("phrase_accuracy", "box_accuracy"):
            code = "from __future__ import division\n" + ast.get_source_segment(src, node)
            exec(compile(code, REF, "exec"), ns)
    return ns["phrase_accuracy"], ns["box_accuracy"]


def make_case(seed, n_imgs, n_cls, dup_labels=False, shuffle=False, ties=False, trailing_empty=2):
    """Arrays describing recs (ground truth per image) and dets (one row per grounded query)."""
    rs = np.random.RandomState(seed)
    gt_img, gt_cls, gt_box, gt_thr = [], [], [], []
    det_img, det_cls, det_box, det_conf = [], [], [], []
    for i in range(n_imgs):
        k = rs.randint(0, 4)
        cls_here = rs.randint(0, n_cls, size=k)  # the same class may be annotated twice in an image
        boxes = []
        for c in cls_here:
            x1, y1 = rs.randint(0, 150, size=2)
            w, h = rs.randint(10, 70, size=2)
            b = [int(x1), int(y1), int(x1 + w), int(y1 + h)]
            boxes.append(b)
            gt_img.append(i); gt_cls.append(int(c)); gt_box.append(b)
            gt_thr.append(0.5)
        if rs.rand() < 0.15 and i != n_imgs - 1:
            continue  # an image without detections
        nd = rs.randint(1, 6)
        if dup_labels:
            labels = rs.randint(0, n_cls, size=nd)
            if k and nd >= 2:
                labels[:2] = cls_here[0]  # two detections of an annotated class
        else:
            labels = rs.permutation(n_cls)[:nd]
            if k:
                labels[0] = cls_here[0]
                labels = np.array(list(dict.fromkeys(labels.tolist())))
        for j, c in enumerate(labels):
            hit = [b for b, cc in zip(boxes, cls_here) if cc == c]
            if hit and rs.rand() < 0.7:
                b = np.array(hit[rs.randint(len(hit))], np.float32)
                if ties and rs.rand() < 0.5:
                    # IoU exactly 0.5 under the +1 convention: same box, half the rows
                    hgt = b[3] - b[1] + 1
                    if int(hgt) % 2 == 0:
                        b = np.array([b[0], b[1], b[2], b[1] + hgt / 2 - 1], np.float32)
                else:
                    b = b + rs.uniform(-12, 12, size=4).astype(np.float32)
            else:
                x1, y1 = rs.uniform(0, 150, size=2)
                b = np.array([x1, y1, x1 + rs.uniform(8, 80), y1 + rs.uniform(8, 80)], np.float32)
            det_img.append(i); det_cls.append(int(c)); det_box.append(b.astype(np.float32))
            det_conf.append(float(rs.uniform(-1, 1)))
    z = dict(gt_img=np.array(gt_img, np.int64), gt_cls=np.array(gt_cls, np.int64),
             gt_box=np.array(gt_box, np.int64).reshape(-1, 4), gt_thr=np.array(gt_thr, np.float64),
             det_img=np.array(det_img, np.int64), det_cls=np.array(det_cls, np.int64),
             det_box=np.array(det_box, np.float32).reshape(-1, 4), det_conf=np.array(det_conf, np.float64),
             n_recs=np.int64(n_imgs + trailing_empty),
             class_list=np.array(["cls%02d" % c for c in range(n_cls)]))
    if shuffle:
        p = rs.permutation(len(det_img))
        for key in ("det_img", "det_cls", "det_box", "det_conf"):
            z[key] = z[key][p]
    return z


def to_reference_inputs(z):
    """recs / dets exactly as model.py:943-983 and parse_gt (youcook_eval.py:78-110) build them."""
    class_list = [str(c) for c in z["class_list"]]
    recs = [dict(label=[], bbox=[], thr=[], img_ids=[]) for _ in range(int(z["n_recs"]))]
    for i, c, b, t in zip(z["gt_img"], z["gt_cls"], z["gt_box"], z["gt_thr"]):
        r = recs[int(i)]
        r["label"].append(class_list[int(c)])
        r["bbox"].append([int(v) for v in b])
        r["thr"].append(float(t))
        r["img_ids"].append(int(i))
    dets = [[int(i) for i in z["det_img"]], [class_list[int(c)] for c in z["det_cls"]],
            [z["det_box"][k] for k in range(len(z["det_box"]))],      # rows of a float32 array
            [np.float64(v) for v in z["det_conf"]]]
    return recs, dets, class_list


CASES = {
    "wellformed": dict(seed=1, n_imgs=60, n_cls=12),
    "duplicate_labels": dict(seed=2, n_imgs=60, n_cls=6, dup_labels=True),
    "unsorted_img_ids": dict(seed=3, n_imgs=40, n_cls=10, shuffle=True),
    "unsorted_duplicates": dict(seed=4, n_imgs=80, n_cls=5, dup_labels=True, shuffle=True),
    "iou_ties": dict(seed=5, n_imgs=50, n_cls=8, ties=True),
    "single_image": dict(seed=6, n_imgs=1, n_cls=3, trailing_empty=0),
}


def main():
    if not os.path.exists(REF):
        sys.exit("needs the reference tree at /root/reference")
    phrase_accuracy, box_accuracy = reference_functions()
    for name, kw in CASES.items():
        z = make_case(**kw)
        recs, dets, class_list = to_reference_inputs(z)
        out = {}
        for fn_name, fn in (("phrase", phrase_accuracy), ("box", box_accuracy)):
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                macro = fn(recs, dets, class_list)
            out[fn_name + "_macro"] = np.float64(macro)
            out[fn_name + "_printed"] = np.array(buf.getvalue().strip().splitlines())
        np.savez_compressed(os.path.join(HERE, "eval_%s.npz" % name), **z, **out)
        print(name, len(z["det_img"]), "dets", len(z["gt_img"]), "gts",
              {k: (float(v) if v.ndim == 0 else list(v)) for k, v in out.items()})


if __name__ == "__main__":
    main()
