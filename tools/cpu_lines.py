"""CPU lines of SURVEY.md section 8(d), run on the GPU box's host cores (reported baselines only):
  (1) reference path: the oracle port (restated NMS tail + RoIAlignAvg in C/OpenMP, torch-CPU DVSA
      fwd(+bwd)) at cfg1 and cfg2 -- what `bench.py --impl reference` times;
  (2) library CPU, NOT parity-equivalent (torchvision's nms / roi_align are different functions,
      SURVEY facts 0.2 / 0.3): torchvision.ops.nms + torchvision.ops.roi_align + the same DVSA.
    python tools/cpu_lines.py [steps]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nafae_b200 import synth  # noqa: E402
from oracle import cpu as ocpu  # noqa: E402
from oracle import dvsa as odvsa  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
cores = os.cpu_count() or 1
torch.set_num_threads(cores)
ocpu.set_num_threads(cores)


def dvsa(b, c):
    if c["train"]:
        return odvsa.dvsa_forward_backward(b["vis_feats"], b["word_feats"], b["lens"], c["Na"], c["Nb"], c["Ne"],
                                           c["Delta"], c["vis_lam"], "train")
    with torch.no_grad():
        return odvsa.dvsa_forward(torch.from_numpy(b["vis_feats"]), torch.from_numpy(b["word_feats"]), b["lens"],
                                  c["Na"], c["Nb"], c["Ne"], c["Delta"], c["vis_lam"], "eval")


def ref_step(b, c):
    rois, _, _ = ocpu.proposal_tail(b["proposals"], b["scores"], c["pre"], c["Nb"], 0.7)
    ocpu.roi_align_avg_forward(b["features"], rois.reshape(-1, 5), 7, 7, 1.0 / 16.0)
    dvsa(b, c)


def lib_step(b, c):
    import torchvision
    F = c["Na"] * c["Ns"]
    feats = torch.from_numpy(b["features"])
    rois = []
    for f in range(F):
        boxes = torch.from_numpy(b["proposals"][f])
        keep = torchvision.ops.nms(boxes, torch.from_numpy(b["scores"][f]), 0.7)[: c["Nb"]]
        rois.append(torch.cat([torch.full((len(keep), 1), float(f)), boxes[keep]], 1))
    torchvision.ops.roi_align(feats, torch.cat(rois, 0), (7, 7), 1.0 / 16.0, sampling_ratio=2, aligned=False)
    dvsa(b, c)


for cfg in ("cfg1", "cfg2"):
    c = synth.CONFIGS[cfg]
    b = synth.make_batch(cfg, 1234)
    for label, fn in (("reference path (oracle port)", ref_step), ("library CPU (torchvision; NOT parity-equivalent)", lib_step)):
        fn(b, c)
        t0 = time.perf_counter()
        for _ in range(steps):
            fn(b, c)
        dt = (time.perf_counter() - t0) / steps
        print("%s %-52s %8.1f ms/step %8.1f segments/s  (%d cores, %s phase)" % (
            cfg, label, dt * 1e3, c["Na"] / dt, cores, "train fwd+bwd" if c["train"] else "eval fwd"), flush=True)
