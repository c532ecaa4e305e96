"""Generates tests/golden/bridge_ebd.npz by executing the reference's own VisEbd / WordEbd class
definitions (model.py:616-642, cut out with `ast`, unchanged) with seeded weights in eval and train
(dropout 0) mode.  Needs /root/reference; run in the build container."""
import ast
import os
import types

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/model.py"


def reference_classes():
    src = open(REF).read()
    ns = {"torch": torch, "nn": nn}
    for node in ast.parse(src).body:
        if isinstance(node, ast.ClassDef) and node.name in ("VisEbd", "WordEbd"):
            exec(compile(ast.get_source_segment(src, node), REF, "exec"), ns)
    return ns["VisEbd"], ns["WordEbd"]


def main():
    VisEbd, WordEbd = reference_classes()
    args = types.SimpleNamespace(vis_fc_dim=64, glove_dim=20, word_ebd_dim=16, dropout_rate=0.0)
    torch.manual_seed(7)
    vis, word = VisEbd(args), WordEbd(args)
    x_vis = torch.randn(12, 64) * 30
    x_word = torch.randn(9, 20)
    out = {}
    for k, v in vis.state_dict().items():
        out["vis_ebd." + k] = v.numpy().copy()
    for k, v in word.state_dict().items():
        out["word_ebd." + k] = v.numpy().copy()
    word.train()
    out["word_train"] = word(x_word).detach().numpy()          # batch statistics
    for k in ("running_mean", "running_var", "num_batches_tracked"):
        out["word_after." + k] = word.state_dict()["bn." + k].numpy().copy()
    vis.eval(); word.eval()
    out["vis_eval"] = vis(x_vis).detach().numpy()
    out["word_eval"] = word(x_word).detach().numpy()            # running statistics
    np.savez_compressed(os.path.join(HERE, "bridge_ebd.npz"), x_vis=x_vis.numpy(), x_word=x_word.numpy(), **out)
    print("wrote bridge_ebd.npz", sorted(out))


if __name__ == "__main__":
    main()
