"""The trainable embedding layers that sit between the two halves of the hot path (SURVEY.md
section 1 "bridge"): plain PyTorch, cuBLAS underneath -- they are NOT part of the kernel work of this
package and exist so that a reference checkpoint's `vis_ebd.*` / `word_ebd.*` tensors load by name
and a complete grounding head (`VisEbd` -> `DVSA` <- `WordEbd`) can be assembled around the
kernels.  Same constructor argument (`args` with `vis_fc_dim`, `glove_dim`, `word_ebd_dim`,
`dropout_rate`), same parameter names, same arithmetic as reference model.py:616-642.
"""
import torch
from torch import nn


class VisEbd(nn.Module):
    """model.py:616-629: RoI fc7 features (R, vis_fc_dim) -> tanh(drop(fc1(x / 100)))."""

    def __init__(self, args):
        super(VisEbd, self).__init__()
        self.fc1 = nn.Linear(args.vis_fc_dim, args.word_ebd_dim)
        self.drop = nn.Dropout(p=args.dropout_rate)

    def forward(self, feats):
        return torch.tanh(self.drop(self.fc1(feats / 100)))


class WordEbd(nn.Module):
    """model.py:631-642: GloVe vectors (Na*Ne, glove_dim) -> tanh(drop(bn(fc1(x)))).  BatchNorm
    statistics stay per replica under data parallelism (the reference is single-GPU; no SyncBN)."""

    def __init__(self, args):
        super(WordEbd, self).__init__()
        self.fc1 = nn.Linear(args.glove_dim, args.word_ebd_dim)
        self.drop = nn.Dropout(p=args.dropout_rate)
        self.bn = nn.BatchNorm1d(args.word_ebd_dim)

    def forward(self, feats):
        return torch.tanh(self.drop(self.bn(self.fc1(feats))))
