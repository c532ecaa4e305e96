"""Inference sweep over many independent segments (reference validate(), model.py:800-991, with
batch_size_val = 1): per segment

    proposal tail (NMS 0.7 -> top-Nb)  ->  RoIAlignAvg 7x7  ->  [bridge: caller's PyTorch]  ->
    DVSA eval forward  ->  postprocess + record_det  ->  box / query accuracy bookkeeping

`EvalStep` runs G segments per launch set: the detector-side kernels batch their frames, the
scoring kernel runs the G segments as independent groups (`nafae_ground_forward_batched`: every
segment sees only its own queries, exactly like G calls with Na = 1), and ONE kernel
(`nafae_eval_record`) turns the picks into recorded detections and per-class accuracy counters on
the device -- no per-step D2H copy, no Python loops over frames and entities (model.py:457-487,
lib/datasets/youcook_eval.py:241-336).  Inputs are passed per call (views into whatever the caller
keeps resident), outputs go to caller-provided slices, so a sweep never copies inside the loop.
"""
import torch

from . import _C


class EvalStep(object):
    KERNELS_PER_STEP = 4  # proposal_tail, align_pool_fwd_slab, ground_fwd, eval_record

    def __init__(self, G, Ns, Nb, Ne, D, C, H, W, n_props, num_classes, pre_nms_topn=6000,
                 nms_thresh=0.7, spatial_scale=1.0 / 16.0, Delta=5.0, vis_lam=1.0, gt_thr=0.5,
                 device=None):
        self.dev = torch.device(device if device is not None else
                                "cuda:%d" % torch.cuda.current_device())
        self.G, self.Ns, self.Nb, self.Ne, self.D = int(G), int(Ns), int(Nb), int(Ne), int(D)
        self.F, self.R = self.G * self.Ns, self.G * self.Ns * self.Nb
        self.C, self.H, self.W, self.n = int(C), int(H), int(W), int(n_props)
        self.pre, self.thresh, self.scale = int(pre_nms_topn), float(nms_thresh), float(spatial_scale)
        self.Delta, self.vis_lam, self.gt_thr = float(Delta), float(vis_lam), float(gt_thr)
        self.num_classes = int(num_classes)
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.rois = torch.empty((self.F, self.Nb, 5), **f32)
        self.roi_scores = torch.empty((self.F, self.Nb), **f32)
        self.pooled = torch.empty((self.R, self.C, 7, 7), **f32)
        self.D_ind = torch.empty((self.G, self.Ns, self.Ne), dtype=torch.int64, device=self.dev)
        self.D_sim = torch.empty((self.G, self.Ns, self.Ne), **f32)
        self.loss = torch.zeros((self.G,), **f32)
        self.ws_group = int(_C.lib.nafae_ground_workspace_bytes(1, self.Ns, self.Nb, self.Ne, self.D))
        self.ws = torch.zeros((self.ws_group * self.G // 4,), dtype=torch.int32, device=self.dev)
        self.align_ws_bytes = int(_C.lib.nafae_roi_align_workspace_bytes(self.F, self.R))
        self.align_ws = torch.zeros((self.align_ws_bytes // 4,), dtype=torch.int32, device=self.dev)

    def run(self, features, proposals, scores, vis_feats, word_feats, lens, image_id_base, out=None,
            gt_boxes=None, gt_classes=None, class_match=None, class_count=None, run_detector=True):
        """features (F, C, H, W); proposals (F, n, 4) / scores (F, n) score-descending per frame;
        vis_feats (G*Ns*Nb, D); word_feats (G*Ne, D); lens (G) int32 -- all CUDA, contiguous.
        out: dict of (G, Ns, Ne[, 4]) tensors `image_ids` int64, `box_rows` int64, `boxes` f32,
        `confs` f32 (any subset).  gt_boxes (G, Ns, Ne, 4) f64 + gt_classes (G, Ne) int32 +
        class_match / class_count (num_classes) int32: accuracy counters, accumulated."""
        L, P = _C.lib, _C.ptr
        s = _C.stream(self.dev)
        out = out or {}
        with torch.cuda.device(self.dev):
            if run_detector:
                _C.check(L.nafae_proposal_tail(P(proposals), P(scores), self.F, self.n, self.pre, self.Nb,
                                               self.thresh, P(self.rois), P(self.roi_scores), None, s),
                         "nafae_proposal_tail")
                _C.check(L.nafae_roi_align_forward(P(features), self.scale, self.F, self.R, self.H, self.W,
                                                   self.C, 7, 7, _C.POOL_AVG, P(self.rois), P(self.pooled),
                                                   _C.FLAG_NO_GATE, P(self.align_ws), self.align_ws_bytes, s),
                         "nafae_roi_align_forward")
            _C.check(L.nafae_ground_forward_batched(P(vis_feats), P(word_feats), P(lens), self.G, 1, self.Ns,
                                                    self.Nb, self.Ne, self.D, self.Delta, self.vis_lam, 0,
                                                    P(self.D_ind), P(self.D_sim), P(self.loss), P(self.ws),
                                                    self.ws.numel() * 4, s), "nafae_ground_forward_batched")
            _C.check(L.nafae_eval_record(P(self.D_ind), P(self.D_sim), P(lens), P(self.rois), self.G, self.Ns,
                                         self.Nb, self.Ne, int(image_id_base), P(out.get("image_ids")),
                                         P(out.get("box_rows")), P(out.get("boxes")), P(out.get("confs")),
                                         P(gt_boxes), P(gt_classes), self.gt_thr, self.num_classes,
                                         P(class_match), P(class_count), s), "nafae_eval_record")


def accuracy_from_counts(class_match, class_count):
    """The reduction box_accuracy / phrase_accuracy end with (youcook_eval.py:228-229, 327-328):
    macro = mean over ALL classes of match / (count + 1e-6), micro = sum(match) / sum(count)."""
    m = class_match.to(torch.float64)
    c = class_count.to(torch.float64)
    return float((m / (c + 1e-6)).mean()), float(m.sum() / c.sum())


def dets_from_records(image_ids, boxes, confs, labels_of):
    """Host lists `[img_ids, labels, boxes, confs]` in the reference's record_det order from the dense
    per-slot tensors EvalStep fills (slots with image id -1 are padded entity slots).
    labels_of(segment_index, entity_index) -> label."""
    ids = image_ids.reshape(-1).cpu().numpy()
    bx = boxes.reshape(-1, 4).cpu().numpy()
    cf = confs.reshape(-1).cpu().numpy()
    S, Ns, Ne = image_ids.shape
    out = ([], [], [], [])
    for i in range(ids.shape[0]):
        if ids[i] < 0:
            continue
        seg, e = i // (Ns * Ne), i % Ne
        out[0].append(int(ids[i]))
        out[1].append(labels_of(seg, e))
        out[2].append(bx[i])
        out[3].append(cf[i])
    return [list(x) for x in out]
