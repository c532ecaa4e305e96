"""Box decoding / clipping used by the proposal layer -- mirror of reference
lib/model/rpn/bbox_transform.py:77-103 (bbox_transform_inv) and :125-133 (clip_boxes).
Plain torch ops on whatever device the inputs live on (these run on the GPU in the reference too)."""
import torch


def bbox_transform_inv(boxes, deltas, batch_size=None):
    """boxes (B, N, 4) anchors, deltas (B, N, 4) = (dx, dy, dw, dh) -> predicted boxes (B, N, 4)."""
    w = boxes[:, :, 2] - boxes[:, :, 0] + 1.0
    h = boxes[:, :, 3] - boxes[:, :, 1] + 1.0
    cx = boxes[:, :, 0] + 0.5 * w
    cy = boxes[:, :, 1] + 0.5 * h
    dx, dy, dw, dh = deltas[:, :, 0::4], deltas[:, :, 1::4], deltas[:, :, 2::4], deltas[:, :, 3::4]
    pcx = dx * w.unsqueeze(2) + cx.unsqueeze(2)
    pcy = dy * h.unsqueeze(2) + cy.unsqueeze(2)
    pw = torch.exp(dw) * w.unsqueeze(2)
    ph = torch.exp(dh) * h.unsqueeze(2)
    out = deltas.clone()
    out[:, :, 0::4] = pcx - 0.5 * pw
    out[:, :, 1::4] = pcy - 0.5 * ph
    out[:, :, 2::4] = pcx + 0.5 * pw
    out[:, :, 3::4] = pcy + 0.5 * ph
    return out


def clip_boxes(boxes, im_shape, batch_size=None):
    """Clamp to [0, W-1] x [0, H-1] per image; im_shape rows are (H, W, scale).  In place."""
    n = boxes.size(0)
    for i in range(n):
        boxes[i, :, 0::4].clamp_(0, float(im_shape[i, 1]) - 1)
        boxes[i, :, 1::4].clamp_(0, float(im_shape[i, 0]) - 1)
        boxes[i, :, 2::4].clamp_(0, float(im_shape[i, 1]) - 1)
        boxes[i, :, 3::4].clamp_(0, float(im_shape[i, 0]) - 1)
    return boxes
