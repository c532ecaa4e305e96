"""Quick per-operator device timings (CUDA events) at a named synthetic config. Dev tool."""
import argparse
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nafae_b200 import synth, _C  # noqa: E402
from nafae_b200.model.rpn.proposal_layer import proposal_tail  # noqa: E402
from nafae_b200.model.nms.nms_wrapper import nms_batched  # noqa: E402
from nafae_b200.model.roi_align.modules.roi_align import RoIAlignAvg  # noqa: E402


def timeit(fn, iters=50, warm=5, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts = np.array(ts)
    return float(np.median(ts)), float(ts.min())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="cfg2")
    ap.add_argument("--exact", action="store_true")
    a = ap.parse_args()
    c = synth.CONFIGS[a.cfg]
    dev = torch.device("cuda:0")
    b = synth.make_batch(a.cfg, 1234)
    feat = torch.from_numpy(b["features"]).to(dev)
    props = torch.from_numpy(b["proposals"]).to(dev)
    scores = torch.from_numpy(b["scores"]).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    F = c["Na"] * c["Ns"]
    rois, _ = proposal_tail(props, scores, c["pre"], c["Nb"], 0.7)
    rois2 = rois.view(-1, 5)
    mod = RoIAlignAvg(7, 7, 1 / 16., exact=a.exact)
    out_bytes = rois2.shape[0] * c["C"] * 49 * 4
    in_bytes = feat.numel() * 4
    featg = feat.clone().requires_grad_(True)
    outg = mod(featg, rois2)
    gy = torch.randn_like(outg)

    def bwd():
        featg.grad = None
        outg.backward(gy, retain_graph=True)
    gin = torch.empty_like(feat)

    def raw_bwd(flags, zero):
        def fn():
            if zero:
                gin.zero_()
            st = _C.lib.nafae_roi_align_backward(_C.ptr(gy), None, 1 / 16., F, rois2.shape[0], c["H"], c["W"], c["C"],
                                                 7, 7, _C.POOL_AVG, _C.ptr(rois2), _C.ptr(gin), flags, _C.stream())
            assert st == 1
        return fn
    bwd_atomic = raw_bwd(_C.FLAG_EXACT, True)   # the reference's formulation: zero-fill + one global atomicAdd per tap
    bwd_scatter = raw_bwd(_C.FLAG_OVERWRITE, False)  # default: shared-memory slab scatter, bulk store
    bwd_scatter_acc = raw_bwd(0, True)          # same, reduce-added onto a zero-filled tensor
    bwd_gather = raw_bwd(_C.FLAG_OVERWRITE | _C.FLAG_DETERMINISTIC, False)  # cell-gather, fixed summation order
    from nafae_b200.model.roi_pooling.modules.roi_pool import _RoIPooling
    pool = _RoIPooling(7, 7, 1 / 16.)
    featp = feat.clone().requires_grad_(True)
    outp = pool(featp, rois2)

    def pool_bwd():
        featp.grad = None
        outp.backward(gy, retain_graph=True)
    for name, fn, nbytes in (
        ("proposal_tail", lambda: proposal_tail(props, scores, c["pre"], c["Nb"], 0.7), None),
        ("nms_batched(full)", lambda: nms_batched(torch.cat((props, scores.unsqueeze(2)), 2), 0.7), None),
        ("roi_align_avg", lambda: mod(feat, rois2), in_bytes + out_bytes),
        ("roi_align_avg bwd (module)", bwd, in_bytes + out_bytes),
        ("roi_align_avg bwd scatter", bwd_scatter, in_bytes + out_bytes),
        ("roi_align_avg bwd zero+scatter-add", bwd_scatter_acc, in_bytes + out_bytes),
        ("roi_align_avg bwd gather (determ.)", bwd_gather, in_bytes + out_bytes),
        ("roi_align_avg bwd zero+global atomics", bwd_atomic, in_bytes + out_bytes),
        ("roi_pool fwd", lambda: pool(feat, rois2), in_bytes + out_bytes),
        ("roi_pool bwd (module)", pool_bwd, in_bytes + out_bytes),
    ):
        med, mn = timeit(fn, flush=flush)
        extra = ""
        if nbytes:
            extra = "  %.0f GB/s (median) %.0f GB/s (best), alg bytes %.1f MB" % (
                nbytes / med / 1e3, nbytes / mn / 1e3, nbytes / 1e6)
        print("%-38s median %8.1f us  min %8.1f us%s" % (name, med, mn, extra))


if __name__ == "__main__":
    main()
