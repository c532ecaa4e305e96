"""Training-step wrapper of the grounding head (reference model.py:762-774):

    optimizer.zero_grad()
    D, D_sim, margin_loss = ground_model.DVSA(vis_feats, word_feats, entities_length)
    loss = criterion(margin_loss, torch.zeros_like(margin_loss))      # nn.L1Loss
    loss.backward()
    torch.nn.utils.clip_grad_norm_(ground_model.parameters(), args.clip)
    optimizer.step()                                                  # Adam, model.py:1077-1082

re-arranged for one process per GPU: the trainable parameters (``vis_ebd.*``, ``word_ebd.*``; the
reference's ``DVSA.*`` parameters never receive a gradient) live in ONE flat fp32 buffer, their
gradients in the flat bucket the data-parallel all-reduce averages, and

    backward  ->  all-reduce (AVG, csrc/allreduce.cu)  ->  clip + Adam (csrc/optim.cu)

so that clipping acts on the AVERAGED gradient and every replica applies the same update
("clip after the all-reduce to keep replicas identical").  The embedding layers
themselves are the caller's PyTorch modules (the "bridge", out of the kernel scope); the scoring /
loss head and everything after ``backward()`` run in this package's kernels.  No CPU path.
"""
import torch

from . import _C
from .grounding import _WorkspacePool, ground


class HeadTrainer(object):
    """vis_ebd / word_ebd: the modules of `nafae_b200.bridge` (or the reference's own, same
    parameter names).  allreduce: a `parallel.PeerAllReduce` / `MulticastAllReduce` sized
    `HeadTrainer.numel(vis_ebd, word_ebd)`, or None for single-GPU training."""

    def __init__(self, vis_ebd, word_ebd, Na, Nb, Ne, Delta, vis_lam, lr=1e-3, weight_decay=1e-5,
                 clip=100.0, betas=(0.9, 0.999), eps=1e-8, allreduce=None):
        self.vis_ebd, self.word_ebd = vis_ebd, word_ebd
        self.Na, self.Nb, self.Ne = int(Na), int(Nb), int(Ne)
        self.Delta, self.vis_lam = float(Delta), float(vis_lam)
        self.lr, self.wd, self.clip = float(lr), float(weight_decay), float(clip)
        self.betas, self.eps = (float(betas[0]), float(betas[1])), float(eps)
        self.allreduce = allreduce
        # parameter order of the reference's optimizer (model.py:1077-1081): word_ebd, then vis_ebd
        self.params = [p for p in list(word_ebd.parameters()) + list(vis_ebd.parameters()) if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        _C.require_cuda(self.params[0], "parameters")
        self.dev = dev
        n = sum(p.numel() for p in self.params)
        self.n = n
        if allreduce is not None:
            if allreduce.buf.numel() < n:
                raise ValueError("all-reduce bucket holds %d floats, parameters need %d"
                                 % (allreduce.buf.numel(), n))
            self.flat_grad = allreduce.buf[:n]
        else:
            self.flat_grad = torch.zeros((n,), dtype=torch.float32, device=dev)
        self.flat_param = torch.empty((n,), dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros((n,), dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros((n,), dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:  # parameters and their .grad become views of the flat buffers
            k = p.numel()
            self.flat_param[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat_param[off:off + k].view_as(p.data)
            p.grad = self.flat_grad[off:off + k].view_as(p.data)
            off += k
        nbytes = int(_C.lib.nafae_clip_adam_workspace_bytes())
        self.opt_ws = torch.zeros((nbytes // 8,), dtype=torch.int64, device=dev)
        self._pool = _WorkspacePool()

    @staticmethod
    def numel(vis_ebd, word_ebd):
        return sum(p.numel() for p in list(word_ebd.parameters()) + list(vis_ebd.parameters())
                   if p.requires_grad)

    def set_lr(self, lr):
        """adjust_learning_rate (model.py:1084-1088)."""
        self.lr = float(lr)

    def forward_backward(self, fc_feats, glove_feats, entities_length):
        """zero_grad, embeddings, DVSA, L1Loss, backward: leaves this rank's gradient in the bucket."""
        self.flat_grad.zero_()
        vis_feats = self.vis_ebd(fc_feats)
        word_feats = self.word_ebd(glove_feats)
        D_ind, D_sim, margin_loss = ground(vis_feats, word_feats, entities_length, self.Na, self.Nb,
                                           self.Ne, self.Delta, self.vis_lam, True, self._pool)
        loss = torch.nn.functional.l1_loss(margin_loss, torch.zeros_like(margin_loss))
        loss.backward()
        return D_ind, D_sim, loss.detach()

    def reduce_and_update(self):
        """all-reduce (AVG) -> clip_grad_norm_ -> Adam, all on the current stream."""
        if self.allreduce is not None and self.allreduce.world > 1:
            self.allreduce.launch()
        P = _C.ptr
        with torch.cuda.device(self.dev):
            st = _C.lib.nafae_clip_adam_step(P(self.flat_param), P(self.flat_grad), P(self.exp_avg),
                                             P(self.exp_avg_sq), self.n, self.lr, self.betas[0],
                                             self.betas[1], self.eps, self.wd, self.clip,
                                             P(self.opt_ws), self.opt_ws.numel() * 8, _C.stream(self.dev))
        _C.check(st, "nafae_clip_adam_step")

    def step(self, fc_feats, glove_feats, entities_length):
        out = self.forward_backward(fc_feats, glove_feats, entities_length)
        self.reduce_and_update()
        return out

    def grad_norm(self):
        """Total gradient norm the last update clipped against (device scalar, no sync)."""
        return self.opt_ws.view(torch.float32)[2]
