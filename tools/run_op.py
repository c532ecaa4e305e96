"""Run one operator a few times (for ncu). Dev tool."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nafae_b200 import synth
from nafae_b200.model.rpn.proposal_layer import proposal_tail
from nafae_b200.model.roi_align.modules.roi_align import RoIAlignAvg

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
c = synth.CONFIGS[cfg]
dev = torch.device("cuda:0")
b = synth.make_batch(cfg, 1234)
feat = torch.from_numpy(b["features"]).to(dev)
props = torch.from_numpy(b["proposals"]).to(dev)
scores = torch.from_numpy(b["scores"]).to(dev)
mod = RoIAlignAvg(7, 7, 1 / 16.)
for _ in range(n):
    rois, _ = proposal_tail(props, scores, c["pre"], c["Nb"], 0.7)
    out = mod(feat, rois.view(-1, 5))
torch.cuda.synchronize()
print("ok", out.shape)
