"""Anchor (reference window) enumeration -- mirror of reference lib/model/rpn/generate_anchors.py:45-105.

Aspect ratios are applied to a `base_size` square keeping its area (widths rounded to integers,
heights = round(width * ratio)), then every ratio anchor is scaled about its centre.  The reference
ships one golden vector for this function (the 9 anchors of base 16, ratios .5/1/2, scales 8/16/32,
generate_anchors.py:12-37); tests/test_proposal_front_cpu.py pins it.
"""
import numpy as np


def _centre_form(box):
    w = box[2] - box[0] + 1.0
    h = box[3] - box[1] + 1.0
    return w, h, box[0] + 0.5 * (w - 1.0), box[1] + 0.5 * (h - 1.0)


def _corner_form(ws, hs, cx, cy):
    ws = np.asarray(ws, dtype=np.float64).reshape(-1, 1)
    hs = np.asarray(hs, dtype=np.float64).reshape(-1, 1)
    return np.hstack((cx - 0.5 * (ws - 1.0), cy - 0.5 * (hs - 1.0),
                      cx + 0.5 * (ws - 1.0), cy + 0.5 * (hs - 1.0)))


def generate_anchors(base_size=16, ratios=(0.5, 1, 2), scales=2 ** np.arange(3, 6)):
    """(len(ratios)*len(scales), 4) float64 anchors [x1, y1, x2, y2] around a (0,0,base-1,base-1) window,
    ratio-major order like the reference."""
    ratios = np.asarray(ratios, dtype=np.float64)
    scales = np.asarray(scales, dtype=np.float64)
    w, h, cx, cy = _centre_form(np.array([0.0, 0.0, base_size - 1.0, base_size - 1.0]))
    ws = np.round(np.sqrt(w * h / ratios))
    hs = np.round(ws * ratios)
    out = []
    for rw, rh in zip(ws, hs):
        out.append(_corner_form(rw * scales, rh * scales, cx, cy))
    return np.vstack(out)
