"""The trainable embedding layers that sit between the two halves of the hot path (SURVEY.md
section 1 "bridge"): plain PyTorch, cuBLAS underneath -- they are NOT part of the kernel work of this
package and exist so that a reference checkpoint's `vis_ebd.*` / `word_ebd.*` tensors load by name
and a complete grounding head (`VisEbd` -> `DVSA` <- `WordEbd`) can be assembled around the
kernels.  Same constructor argument (`args` with `vis_fc_dim`, `glove_dim`, `word_ebd_dim`,
`dropout_rate`), same parameter names, same arithmetic as reference model.py:616-642.
"""
import torch
from torch import nn


class RCNNTop(nn.Module):
    """Inference-only `RCNN_top` of the frozen detector (lib/model/faster_rcnn/vgg16_rpn.py:35,56-61:
    VGG16 fc6 + ReLU + Dropout, fc7 + ReLU + Dropout in eval mode) on the tcgen05 tensor cores:
    weights are converted to bf16 once, each layer is one `nafae_gemm_bf16_tn` launch (bias + ReLU in
    the epilogue), accumulation is fp32.  `forward(pooled)` takes the RoIAlign output viewed as
    (R, C*7*7) in fp32 (cast here) or bf16 and returns fc7 features (R, out) in fp32 -- the input of
    `VisEbd`.  The detector is frozen in NAFAE (model.py:651,673), so there is no backward."""

    def __init__(self, fc6, fc7):
        super(RCNNTop, self).__init__()
        self.w6 = nn.Parameter(fc6.weight.detach().to(torch.bfloat16).contiguous(), requires_grad=False)
        self.b6 = nn.Parameter(fc6.bias.detach().float().contiguous(), requires_grad=False)
        self.w7 = nn.Parameter(fc7.weight.detach().to(torch.bfloat16).contiguous(), requires_grad=False)
        self.b7 = nn.Parameter(fc7.bias.detach().float().contiguous(), requires_grad=False)

    @torch.no_grad()
    def forward(self, pooled):
        from . import _C
        _C.require_cuda(pooled, "pooled")
        x = pooled.reshape(pooled.shape[0], -1)
        if x.dtype != torch.bfloat16:
            x = x.to(torch.bfloat16)
        x = x.contiguous()
        R = x.shape[0]
        h = torch.empty((R, self.w6.shape[0]), dtype=torch.bfloat16, device=x.device)
        y = torch.empty((R, self.w7.shape[0]), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            s = _C.stream(x.device)
            _C.check(_C.lib.nafae_gemm_bf16_tn(_C.ptr(x), _C.ptr(self.w6), _C.ptr(self.b6), _C.ptr(h), R,
                                               self.w6.shape[0], self.w6.shape[1], 3, s), "nafae_gemm_bf16_tn(fc6)")
            _C.check(_C.lib.nafae_gemm_bf16_tn(_C.ptr(h), _C.ptr(self.w7), _C.ptr(self.b7), _C.ptr(y), R,
                                               self.w7.shape[0], self.w7.shape[1], 1, s), "nafae_gemm_bf16_tn(fc7)")
        return y


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
