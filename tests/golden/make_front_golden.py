"""Generates tests/golden/front_*.npz by EXECUTING the reference's own proposal front end on seeded RPN
outputs: `generate_anchors` and its helpers (lib/model/rpn/generate_anchors.py:45-105),
`bbox_transform_inv` and `clip_boxes` (lib/model/rpn/bbox_transform.py:77-103,125-133) are cut out of
the source files with `ast` and exec'd unchanged; the surrounding lines of `_ProposalLayer.forward`
(proposal_layer.py:66-125: anchor shifts, NCHW -> (H, W, A) re-ordering, score sort) are restated here
line by line because that method cannot run without the compiled NMS extension.
Run in the build container (needs /root/reference); the fixtures travel.

    python tests/golden/make_front_golden.py
"""
import ast
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_RPN = "/root/reference/lib/model/rpn"


def _functions(path, names, ns):
    src = open(path).read()
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.get_source_segment(src, node), path, "exec"), ns)
    return ns


def reference_functions():
    ns = {"np": np, "torch": torch, "xrange": range}  # the one Python-2 name generate_anchors.py uses
    _functions(os.path.join(REF_RPN, "generate_anchors.py"),
               ("generate_anchors", "_whctrs", "_mkanchors", "_ratio_enum", "_scale_enum"), ns)
    _functions(os.path.join(REF_RPN, "bbox_transform.py"), ("bbox_transform_inv", "clip_boxes"), ns)
    return ns


def front(ns, cls_prob, deltas, im_info, feat_stride, scales, ratios):
    """proposal_layer.py:66-125 with the reference's own helper functions."""
    anchors0 = torch.from_numpy(ns["generate_anchors"](scales=np.array(scales), ratios=np.array(ratios))).float()
    A = anchors0.size(0)
    scores = cls_prob[:, A:, :, :]                                                        # :66
    batch_size = deltas.size(0)
    feat_height, feat_width = scores.size(2), scores.size(3)
    shift_x = np.arange(0, feat_width) * feat_stride                                      # :80-84
    shift_y = np.arange(0, feat_height) * feat_stride
    shift_x, shift_y = np.meshgrid(shift_x, shift_y)
    shifts = torch.from_numpy(np.vstack((shift_x.ravel(), shift_y.ravel(),
                                         shift_x.ravel(), shift_y.ravel())).transpose())
    shifts = shifts.contiguous().type_as(scores).float()
    K = shifts.size(0)
    anchors = anchors0.view(1, A, 4) + shifts.view(K, 1, 4)                               # :92
    anchors = anchors.view(1, K * A, 4).expand(batch_size, K * A, 4)
    bbox_deltas = deltas.permute(0, 2, 3, 1).contiguous().view(batch_size, -1, 4)         # :98-99
    scores = scores.permute(0, 2, 3, 1).contiguous().view(batch_size, -1)                 # :102-103
    proposals = ns["bbox_transform_inv"](anchors, bbox_deltas, batch_size)                # :106
    proposals = ns["clip_boxes"](proposals, im_info, batch_size)                          # :109
    # :125 torch.sort(scores_keep, 1, True): the order of equal scores is unspecified there; the
    # fixture pins the stable one (ties keep ascending anchor index), which is what the kernel produces
    _, order = torch.sort(scores, stable=True, dim=1, descending=True)
    return anchors0, proposals, scores, order


def make_case(seed, B, H, W, img_h, img_w, ties):
    """Seeded RPN outputs made of exactly representable values only (integer draws scaled by powers of
    two), so that every machine regenerates bit-identical inputs -- softmax / randn differ by an ulp
    between CPU builds, which would reorder near-equal scores."""
    g = torch.Generator().manual_seed(seed)
    A = 12
    fg = torch.randint(1, 1 << 20, (B, A, H, W), generator=g).float() / float(1 << 20)     # (0, 1)
    cls_prob = torch.cat([1.0 - fg, fg], 1)                                               # bg | fg, rpn.py:87-90
    if ties:  # saturated and repeated probabilities, as a confident RPN produces them
        cls_prob[:, A:, :, : W // 2] = torch.round(cls_prob[:, A:, :, : W // 2] * 8) / 8
        cls_prob[:, A:, 0, :] = 1.0
    deltas = torch.randint(-(1 << 15), 1 << 15, (B, 4 * A, H, W), generator=g).float() / float(1 << 16)
    deltas[:, 2::4] *= 0.5
    deltas[:, 3::4] *= 0.5
    im_info = torch.tensor([[float(img_h), float(img_w), 1.0]] * B)
    return cls_prob.contiguous(), deltas.contiguous(), im_info


def main():
    ns = reference_functions()
    for name, seed, B, H, W, ih, iw, ties in (("small", 1, 3, 5, 7, 80, 112, False),
                                              ("real_14x14", 2, 4, 14, 14, 224, 224, False),
                                              ("ties_14x14", 3, 2, 14, 14, 224, 224, True),
                                              ("map_38x50", 4, 2, 38, 50, 608, 800, True)):
        cls_prob, deltas, im_info = make_case(seed, B, H, W, ih, iw, ties)
        anchors, proposals, scores, order = front(ns, cls_prob, deltas, im_info, 16, [4, 8, 16, 32], [0.5, 1, 2])
        big = proposals.shape[1] > 4000  # keep the fixture small: the inputs are regenerated from the seed
        np.savez_compressed(os.path.join(HERE, "front_%s.npz" % name), seed=seed, B=B, H=H, W=W, img_h=ih,
                            img_w=iw, ties=ties, anchors=anchors.numpy(),
                            proposals=proposals.numpy()[:, :: (7 if big else 1)],
                            order=order.numpy().astype(np.int32)[:, : (4000 if big else order.shape[1])],
                            scores_sorted=torch.gather(scores, 1, order).numpy()[:, : (4000 if big else order.shape[1])],
                            stride=7 if big else 1)
        print(name, tuple(proposals.shape), "ties in scores:", int((scores[:, 1:] == scores[:, :-1]).sum()))


if __name__ == "__main__":
    main()
