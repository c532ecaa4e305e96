""""With bridge" line of SURVEY.md section 8(d): the path's kernels with the bridge layers between the
halves -- RCNN_top fc6 / fc7 (frozen; this package's tcgen05 bf16 GEMM vs PyTorch fp32 / bf16 cuBLAS)
and VisEbd / WordEbd (trainable, PyTorch) -- timed per stage with CUDA events at cfg2.
    python tools/time_bridge.py"""
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nafae_b200 import synth  # noqa: E402
from nafae_b200.bridge import RCNNTop, VisEbd, WordEbd  # noqa: E402
from nafae_b200.grounding import ground  # noqa: E402
from nafae_b200.pipeline import GroundingStep  # noqa: E402

dev = torch.device("cuda:0")
c = synth.CONFIGS["cfg2"]
st = GroundingStep(c["Na"], c["Ns"], c["Nb"], c["Ne"], c["D"], c["C"], c["H"], c["W"], c["n"], device=dev)
b = synth.make_batch("cfg2", 1234)
st.load(b)
torch.manual_seed(0)
fc6 = torch.nn.Linear(25088, 4096).to(dev)
fc7 = torch.nn.Linear(4096, 4096).to(dev)
args = types.SimpleNamespace(vis_fc_dim=4096, glove_dim=200, word_ebd_dim=512, dropout_rate=0.1)
vis_ebd, word_ebd = VisEbd(args).to(dev), WordEbd(args).to(dev)
top = RCNNTop(fc6, fc7)
glove = torch.randn(c["Na"] * c["Ne"], 200, device=dev) * 0.4
w6b, w7b = fc6.weight.detach().to(torch.bfloat16), fc7.weight.detach().to(torch.bfloat16)


def timed(fn, n=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def head(fc):
    vis = vis_ebd(fc)
    word = word_ebd(glove)
    D_ind, D_sim, loss = ground(vis, word, b["lens"], c["Na"], c["Nb"], c["Ne"], c["Delta"], c["vis_lam"], True)
    torch.nn.functional.l1_loss(loss, torch.zeros_like(loss)).backward()


st.run_detector()
x32 = st.pooled.view(st.R, -1)
xb = x32.to(torch.bfloat16)
rows = []
rows.append(("detector half (proposal tail + RoIAlignAvg, this package)", timed(st.run_detector)))
rows.append(("pooled fp32 -> bf16 cast (torch)", timed(lambda: x32.to(torch.bfloat16))))
rows.append(("RCNN_top fc6+fc7, this package (tcgen05 bf16, bias+ReLU fused)", timed(lambda: top(xb))))
with torch.no_grad():
    rows.append(("RCNN_top fc6+fc7, PyTorch fp32 (cuBLAS, reference arithmetic)",
                 timed(lambda: torch.relu(fc7(torch.relu(fc6(x32)))), n=5)))
    rows.append(("RCNN_top fc6+fc7, PyTorch bf16 (cuBLAS)",
                 timed(lambda: torch.relu(torch.nn.functional.linear(torch.relu(torch.nn.functional.linear(
                     xb, w6b, fc6.bias.to(torch.bfloat16))), w7b, fc7.bias.to(torch.bfloat16))))))
    fc_feats = top(xb)
    want = torch.relu(fc7(torch.relu(fc6(x32))))
    err = float((fc_feats - want).abs().max() / want.abs().max())
rows.append(("VisEbd / WordEbd + DVSA fwd+bwd + embedding backward (PyTorch + this package)", timed(lambda: head(fc_feats))))
flops = 2.0 * st.R * (25088 * 4096 + 4096 * 4096)
for name, us in rows:
    extra = ""
    if "RCNN_top" in name:
        extra = "  %.0f TFLOP/s" % (flops / us / 1e6)
    print("%-80s %9.1f us%s" % (name, us, extra))
print("fc7 features, tcgen05 bf16 vs fp32 layers: max rel err %.3e" % err)
print("with-bridge step (detector half + fc6/fc7 tcgen05 + head): %.1f us -> %.0f segments/s" % (
    rows[0][1] + rows[1][1] + rows[2][1] + rows[5][1], c["Na"] / ((rows[0][1] + rows[1][1] + rows[2][1] + rows[5][1]) * 1e-6)))
