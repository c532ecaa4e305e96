"""Scoring / loss head: drop-in for the reference ``DVSA`` module, ``postprocess`` and ``record_det``
(reference model.py:457-614) on top of the fused sm_100a kernels (csrc/ground.cu).

    dvsa = DVSA(args, cfg); dvsa.init_train()
    D_ind, D_sim, margin_loss = dvsa(vis_feats, word_feats, entities_length)
    loss = criterion(margin_loss, zeros); loss.backward()          # model.py:768-772

One forward kernel and (train) two backward kernels replace ~60 ATen ops, the numpy mask build
and its H2D copy, Na*Ns*Ne scalar device writes and a ``.nonzero()`` host sync per step.
"""
import numpy as np
import torch

from . import _C


class _Ground(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vis_feats, word_feats, lens, dims, Delta, vis_lam, train, pool, recording, tensor_cores):
        Na, Ns, Nb, Ne, D = dims
        vis = _C.f32c(vis_feats, "vis_feats")
        word = _C.f32c(word_feats, "word_feats")
        dev = vis.device
        F, NQ = Na * Ns, Na * Ne
        if vis.shape != (F * Nb, D) or word.shape != (NQ, D):
            raise ValueError("vis_feats must be (Na*Ns*Nb, D) and word_feats (Na*Ne, D); got %s, %s"
                             % (tuple(vis.shape), tuple(word.shape)))
        D_ind = torch.empty((F, NQ), dtype=torch.int64, device=dev)
        D_sim = torch.empty((F, NQ), dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        ws = pool.take(dims, dev)
        fwd = _C.lib.nafae_ground_forward_tc if tensor_cores else _C.lib.nafae_ground_forward
        with torch.cuda.device(dev):
            st = fwd(_C.ptr(vis), _C.ptr(word), _C.ptr(lens), Na, Ns, Nb, Ne, D, float(Delta), float(vis_lam),
                     int(train), _C.ptr(D_ind), _C.ptr(D_sim), _C.ptr(loss), _C.ptr(ws), ws.numel() * 4,
                     _C.stream(dev))
        _C.check(st, "nafae_ground_forward_tc" if tensor_cores else "nafae_ground_forward")
        # a backward can only follow when a graph is being recorded (`recording` = the caller's
        # torch.is_grad_enabled(); inside Function.forward grad mode is always off): under
        # torch.no_grad() -- the usual validation loop over live modules -- the workspace goes
        # straight back to the pool
        needs_bwd = recording and (vis_feats.requires_grad or word_feats.requires_grad)
        ctx.cfg = (dims, float(Delta), float(vis_lam), int(train))
        ctx.pool = pool
        if needs_bwd:
            ctx.ws = ws
            ctx.save_for_backward(vis, word, lens, D_ind, D_sim)
        else:
            pool.give(dims, dev, ws)
        ctx.mark_non_differentiable(D_ind, D_sim)
        return D_ind, D_sim, loss

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, _g_ind, _g_sim, g_loss):
        (Na, Ns, Nb, Ne, D), Delta, vis_lam, train = ctx.cfg
        vis, word, lens, D_ind, D_sim = ctx.saved_tensors
        ws = ctx.ws
        if ws is None:
            raise RuntimeError("nafae_b200 DVSA: the fused backward consumes the forward's workspace; a "
                               "second backward through the same forward (retain_graph=True) is not "
                               "supported -- run the forward again")
        dev = vis.device
        g_loss = g_loss.to(dtype=torch.float32).contiguous()
        gvis = torch.empty_like(vis)
        gword = torch.empty_like(word)
        with torch.cuda.device(dev):
            st = _C.lib.nafae_ground_backward(_C.ptr(g_loss), _C.ptr(vis), _C.ptr(word),
                                              _C.ptr(lens), Na, Ns, Nb, Ne, D, Delta, vis_lam,
                                              train, _C.ptr(D_ind), _C.ptr(D_sim), _C.ptr(gvis),
                                              _C.ptr(gword), _C.ptr(ws), ws.numel() * 4,
                                              _C.stream(dev))
        _C.check(st, "nafae_ground_backward")
        ctx.pool.give((Na, Ns, Nb, Ne, D), dev, ws)
        ctx.ws = None
        return gvis, gword, None, None, None, None, None, None, None, None


class _WorkspacePool(object):
    """Zero-initialised forward->backward scratch buffers, recycled (the kernels leave their
    counters and accumulators clean on exit, so no memset per step)."""

    def __init__(self):
        self._free = {}

    def take(self, dims, dev):
        lst = self._free.setdefault((dims, str(dev)), [])
        if lst:
            return lst.pop()
        nbytes = int(_C.lib.nafae_ground_workspace_bytes(*dims))
        return torch.zeros((nbytes // 4,), dtype=torch.int32, device=dev)

    def give(self, dims, dev, ws):
        self._free.setdefault((dims, str(dev)), []).append(ws)


def _lens_tensor(entities_length, device):
    if torch.is_tensor(entities_length):
        return entities_length.to(device=device, dtype=torch.int32).contiguous()
    return torch.tensor([int(x) for x in entities_length], dtype=torch.int32, device=device)


def ground(vis_feats, word_feats, entities_length, Na, Nb, Ne, Delta, vis_lam, train, pool=None,
           tensor_cores=False):
    """Functional form.  ``entities_length``: list of ints (as in the reference) or an int32 CUDA
    tensor (no H2D copy).  Returns (D_ind int64 (Na*Ns, Na*Ne), D_sim f32, margin_loss 0-dim).
    ``tensor_cores``: run the regions x queries contraction as tcgen05 tf32x3 tiles
    (`nafae_ground_forward_tc`: same picks, D_sim / loss within 2e-4 relative); "auto" chooses by
    shape from the measured A/B."""
    _C.require_cuda(vis_feats, "vis_feats")
    if tensor_cores == "auto":
        # measured A/B (profiles/RESULTS.md): the tcgen05 form wins once there are hundreds of query
        # columns (520 live columns: 51 us vs 377 us); at the reference's 104 slots the FMA form does
        tensor_cores = int(Na) * int(Ne) >= TC_AUTO_MIN_COLUMNS and int(Nb) <= 128
    Ns = int(vis_feats.size(0) / Na / Nb)  # model.py:530
    dims = (int(Na), Ns, int(Nb), int(Ne), int(vis_feats.size(1)))
    lens = _lens_tensor(entities_length, vis_feats.device)
    if lens.numel() != Na:
        raise ValueError("entities_length must have Na=%d entries" % Na)
    return _Ground.apply(vis_feats, word_feats, lens, dims, Delta, vis_lam, train,
                         pool if pool is not None else _DEFAULT_POOL, torch.is_grad_enabled(),
                         bool(tensor_cores))


_DEFAULT_POOL = _WorkspacePool()
TC_AUTO_MIN_COLUMNS = 256  # query slots (Na * Ne) from which tensor_cores="auto" picks the tcgen05 form


class DVSA(torch.nn.Module):
    """Mirror of reference ``DVSA`` (model.py:490-614): same constructor arguments, ``init_train`` /
    ``init_eval``, ``forward(vis_feats, word_feats, entities_length)``.

    Reads ``args.{batch_size, batch_size_val, max_ent_len, Delta, vis_lam}`` and
    ``cfg.TEST.RPN_POST_NMS_TOP_N`` like the reference.  The reference also constructs
    ``slf_attn`` / ``position_enc`` / ``ffn`` (model.py:495-499) which its forward never uses (dead
    parameters with ``grad=None``); they are not created here -- load reference checkpoints with
    ``strict=False``.
    """

    def __init__(self, args, cfg):
        super(DVSA, self).__init__()
        self.args = args
        self.cfg = cfg
        self.phase = ''
        self.tensor_cores = bool(getattr(args, "tensor_cores", False))  # tcgen05 form of the contraction
        self._pool = _WorkspacePool()

    def init_train(self):
        self.Na = self.args.batch_size
        self.phase = 'train'

    def init_eval(self):
        self.Na = self.args.batch_size_val
        self.phase = 'eval'

    def forward(self, vis_feats, word_feats, entities_length):
        if self.phase not in ('train', 'eval'):
            raise RuntimeError("call init_train() or init_eval() first (model.py:509-515)")
        Nb = self.cfg.TEST.RPN_POST_NMS_TOP_N
        return ground(vis_feats, word_feats, entities_length, self.Na, Nb, self.args.max_ent_len,
                      self.args.Delta, self.args.vis_lam, self.phase == 'train', self._pool,
                      self.tensor_cores)


def postprocess(D, D_sim, Na, Ns, Nb, Ne):
    """model.py:457-474.  Accepts CUDA tensors (runs on the device, returns CUDA tensors
    (Na,Ns,Ne) int64 / f32) or, like the reference, numpy arrays (copied to the device and back)."""
    as_numpy = not torch.is_tensor(D)
    if as_numpy:
        dev = torch.device("cuda", torch.cuda.current_device())
        D = torch.as_tensor(np.ascontiguousarray(D), dtype=torch.int64).to(dev)
        D_sim = torch.as_tensor(np.ascontiguousarray(D_sim), dtype=torch.float32).to(dev)
    _C.require_cuda(D, "D")
    D = D.to(torch.int64).contiguous()
    D_sim = _C.f32c(D_sim, "D_sim")
    out = torch.empty((Na, Ns, Ne), dtype=torch.int64, device=D.device)
    out_sim = torch.empty((Na, Ns, Ne), dtype=torch.float32, device=D.device)
    with torch.cuda.device(D.device):
        st = _C.lib.nafae_ground_postprocess(_C.ptr(D), _C.ptr(D_sim), Na, Ns, Nb, Ne, _C.ptr(out),
                                             _C.ptr(out_sim), _C.stream(D.device))
    _C.check(st, "nafae_ground_postprocess")
    if as_numpy:
        return out.cpu().numpy().astype(int), out_sim.cpu().numpy().astype(np.float64)
    return out, out_sim


def record_det(img_inds, obj_labels, obj_bboxes, obj_confs, Nb, vid_entities, D, D_sim, img_ids,
               infer_boxes):
    """model.py:477-487 (host-side bookkeeping, same in-place list appends)."""
    Na, Ns, Ne = D.shape
    for act_ind, entities in enumerate(vid_entities):
        for spl_ind in range(Ns):
            for ent_ind, entity in enumerate(entities):
                box_id_offset = int(D[act_ind][spl_ind][ent_ind])
                img_inds.append(img_ids[box_id_offset // Nb])
                obj_labels.append(entity)
                obj_bboxes.append(infer_boxes[box_id_offset])
                obj_confs.append(D_sim[act_ind][spl_ind][ent_ind])


def record_det_tensors(D, D_sim, entities_length, img_ids, infer_boxes, Nb):
    """Device-resident, loop-free form of `record_det` (model.py:477-487) for the evaluation sweep:
    every (segment a, frame s, entity e < entities_length[a]) in the reference's append order
    (a, then s, then e) as tensors on D's device.

    D, D_sim: (Na, Ns, Ne) from `postprocess` (global box rows / similarities); img_ids: (F,)
    integer tensor (one id per frame row); infer_boxes: (R, 4).  Returns
    ``(img_inds, seg_idx, ent_idx, boxes, confs)``: `vid_entities[seg_idx[i]][ent_idx[i]]` is the
    label the reference appends for row i.  No host synchronisation except the final sizes."""
    Na, Ns, Ne = D.shape
    dev = D.device
    lens = _lens_tensor(entities_length, dev).to(torch.int64)
    ent = torch.arange(Ne, device=dev)
    keep = (ent.view(1, 1, Ne) < lens.view(Na, 1, 1)).expand(Na, Ns, Ne)
    rows = D[keep].to(torch.int64)                      # row-major: a, s, e -- the reference's order
    seg_idx = torch.arange(Na, device=dev).view(Na, 1, 1).expand(Na, Ns, Ne)[keep]
    ent_idx = ent.view(1, 1, Ne).expand(Na, Ns, Ne)[keep]
    img_inds = torch.as_tensor(img_ids, device=dev)[torch.div(rows, int(Nb), rounding_mode="floor")]
    boxes = torch.as_tensor(infer_boxes, device=dev)[rows]
    return img_inds, seg_idx, ent_idx, boxes, D_sim[keep]
