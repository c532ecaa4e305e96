"""Dev tool: per-kernel device times of one GroundingStep in steady state (sustained clocks).
Events are recorded between the C-ABI calls of a warm loop; medians over many iterations."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nafae_b200 import synth, _C
from nafae_b200.pipeline import GroundingStep

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 300
c = synth.CONFIGS[cfg]
dev = torch.device("cuda:0")
steps = []
for i in range(2):
    st = GroundingStep(c["Na"], c["Ns"], c["Nb"], c["Ne"], c["D"], c["C"], c["H"], c["W"], c["n"],
                       pre_nms_topn=c["pre"], Delta=c["Delta"], vis_lam=c["vis_lam"], train=c["train"], device=dev)
    st.load(synth.make_batch(cfg, 1234 + i))
    steps.append(st)
L, P = _C.lib, _C.ptr
s = _C.stream(dev)

def calls(st):
    Na, Ns, Nb, Ne, D = st.dims
    return [
        ("proposal_tail", lambda: L.nafae_proposal_tail(P(st.proposals), P(st.scores), st.F, st.n, st.pre, Nb, st.thresh, P(st.rois), P(st.roi_scores), None, s)),
        ("roi_align", lambda: L.nafae_roi_align_forward(P(st.features), st.scale, st.F, st.R, st.H, st.W, st.C, 7, 7, 1, P(st.rois), P(st.pooled), 0, None, 0, s)),
        ("ground_fwd", lambda: L.nafae_ground_forward(P(st.vis_feats), P(st.word_feats), P(st.lens), Na, Ns, Nb, Ne, D, st.Delta, st.vis_lam, int(st.train), P(st.D_ind), P(st.D_sim), P(st.loss), P(st.ws), st.ws.numel() * 4, s)),
        ("ground_bwd", lambda: L.nafae_ground_backward(P(st.grad_loss), P(st.vis_feats), P(st.word_feats), P(st.lens), Na, Ns, Nb, Ne, D, st.Delta, st.vis_lam, int(st.train), P(st.D_ind), P(st.D_sim), P(st.grad_vis), P(st.grad_word), P(st.ws), st.ws.numel() * 4, s)),
    ]

cl = [calls(st) for st in steps]
names = [n for n, _ in cl[0]]
for i in range(20):
    for _, fn in cl[i & 1]:
        assert fn() == 1
torch.cuda.synchronize()
times = {n: [] for n in names}
evs = [[torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)] for _ in range(iters)]
for i in range(iters):
    evs[i][0].record()
    for k, (_, fn) in enumerate(cl[i & 1]):
        fn()
        evs[i][k + 1].record()
torch.cuda.synchronize()
for i in range(iters):
    for k, n in enumerate(names):
        times[n].append(evs[i][k].elapsed_time(evs[i][k + 1]) * 1e3)
tot = 0
for n in names:
    t = np.array(times[n][20:])
    print("%-14s median %7.2f us   p10 %7.2f   p90 %7.2f" % (n, np.median(t), np.percentile(t, 10), np.percentile(t, 90)))
    tot += np.median(t)
print("sum of medians %.1f us (includes ~1 us event gaps per kernel)" % tot)
