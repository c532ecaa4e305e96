"""Pre-allocated, graph-capturable execution of the whole hot path for one batch of segments:

    proposal tail (batched NMS + top-N + padding)  ->  RoIAlignAvg 7x7  ->  [bridge: caller's
    PyTorch]  ->  similarity + losses forward  ->  backward (dL/dvis_feats, dL/dword_feats)

``GroundingStep`` is the public fast path: every buffer is allocated once, every kernel is
launched through the C ABI on the current stream with no host synchronisation, so ``run()`` can be
captured into a CUDA graph (``capture()`` / ``replay()``).  The module classes in ``nafae_b200.model``
and ``nafae_b200.grounding`` are the drop-in, autograd-friendly form of the same kernels.
"""
import torch

from . import _C


class GroundingStep(object):
    KERNELS_PER_STEP_TRAIN = 4  # proposal_tail, align_pool_fwd_slab, ground_fwd, ground_bwd
    KERNELS_PER_STEP_EVAL = 3

    def __init__(self, Na, Ns, Nb, Ne, D, C, H, W, n_props, pre_nms_topn=6000, nms_thresh=0.7,
                 spatial_scale=1.0 / 16.0, Delta=10.0, vis_lam=4.13, train=True, device=None,
                 l1_loss=True, tensor_cores=False):
        self.dev = torch.device(device if device is not None else
                                "cuda:%d" % torch.cuda.current_device())
        self.dims = (Na, Ns, Nb, Ne, D)
        self.F, self.R, self.NQ = Na * Ns, Na * Ns * Nb, Na * Ne
        self.C, self.H, self.W, self.n = C, H, W, n_props
        self.pre, self.thresh, self.scale = int(pre_nms_topn), float(nms_thresh), float(spatial_scale)
        self.Delta, self.vis_lam, self.train = float(Delta), float(vis_lam), bool(train)
        self.tensor_cores = bool(tensor_cores)  # contraction of the scoring kernel on tcgen05 (tf32x3)
        f32 = dict(dtype=torch.float32, device=self.dev)
        # device-resident inputs
        self.features = torch.empty((self.F, C, H, W), **f32)
        self.proposals = torch.empty((self.F, n_props, 4), **f32)
        self.scores = torch.empty((self.F, n_props), **f32)
        self.vis_feats = torch.empty((self.R, D), **f32)
        self.word_feats = torch.empty((self.NQ, D), **f32)
        self.lens = torch.zeros((Na,), dtype=torch.int32, device=self.dev)
        # outputs
        self.rois = torch.empty((self.F, Nb, 5), **f32)
        self.roi_scores = torch.empty((self.F, Nb), **f32)
        self.pooled = torch.empty((self.R, C, 7, 7), **f32)
        self.D_ind = torch.empty((self.F, self.NQ), dtype=torch.int64, device=self.dev)
        self.D_sim = torch.empty((self.F, self.NQ), **f32)
        self.loss = torch.zeros((), **f32)
        # upstream gradient of margin_loss.  l1_loss: the step wrapper's L1Loss(margin_loss, 0)
        # (reference model.py:771) -> sign(margin_loss), taken by the backward kernel from the
        # forward's workspace (NULL pointer); otherwise a device scalar the caller may overwrite
        self.grad_loss = None if l1_loss else torch.ones((), **f32)
        self.grad_vis = torch.empty((self.R, D), **f32)
        self.grad_word = torch.empty((self.NQ, D), **f32)
        nbytes = int(_C.lib.nafae_ground_workspace_bytes(*self.dims))
        self.ws = torch.zeros((nbytes // 4,), dtype=torch.int32, device=self.dev)
        # RoIAlign workspace: the persistent kernel's residency gate (include/nafae_b200.h)
        # + one claim counter per frame (dynamic unit scheduling)
        self.align_ws_bytes = int(_C.lib.nafae_roi_align_workspace_bytes(self.F, self.R))
        self.gate = torch.zeros((self.align_ws_bytes // 4,), dtype=torch.int32, device=self.dev)
        self.graph = None

    # -- input staging ---------------------------------------------------------------------
    def load(self, batch, non_blocking=False):
        """Copy a host batch (dict of numpy arrays / CPU tensors as made by synth.make_batch) in."""
        def put(dst, src):
            t = src if torch.is_tensor(src) else torch.from_numpy(src)
            dst.copy_(t.view(dst.shape), non_blocking=non_blocking)
        put(self.features, batch["features"])
        put(self.proposals, batch["proposals"])
        put(self.scores, batch["scores"])
        put(self.vis_feats, batch["vis_feats"])
        put(self.word_feats, batch["word_feats"])
        lens = batch["lens"]
        if not torch.is_tensor(lens):
            lens = torch.tensor([int(x) for x in lens], dtype=torch.int32)
        self.lens.copy_(lens.to(torch.int32), non_blocking=non_blocking)

    def h2d_bytes(self):
        return sum(t.numel() * t.element_size() for t in
                   (self.features, self.proposals, self.scores, self.vis_feats, self.word_feats,
                    self.lens))

    # -- the step --------------------------------------------------------------------------
    def run_tail(self):
        """Proposal tail: batched NMS + top-N + zero padding -> rois, roi_scores."""
        Nb = self.dims[2]
        L, P = _C.lib, _C.ptr
        with torch.cuda.device(self.dev):
            _C.check(L.nafae_proposal_tail(P(self.proposals), P(self.scores), self.F, self.n,
                                           self.pre, Nb, self.thresh, P(self.rois),
                                           P(self.roi_scores), None, _C.stream(self.dev)),
                     "nafae_proposal_tail")

    def run_align(self, gated=True):
        """RoIAlignAvg 7x7 of the current rois -> pooled (R, C, 7, 7).  gated: `self.gate` opens once
        all persistent CTAs of the kernel are resident (concurrent branches may `wait_gate` on it)."""
        L, P = _C.lib, _C.ptr
        with torch.cuda.device(self.dev):
            _C.check(L.nafae_roi_align_forward(P(self.features), self.scale, self.F, self.R, self.H,
                                               self.W, self.C, 7, 7, _C.POOL_AVG, P(self.rois),
                                               P(self.pooled), 0 if gated else _C.FLAG_NO_GATE,
                                               P(self.gate), self.align_ws_bytes, _C.stream(self.dev)),
                     "nafae_roi_align_forward")

    def sync_gate(self):
        """Mark every gated RoIAlign launch so far as seen by all waiting branches (enqueue after
        gated launches that had no `wait_gate` partner, e.g. warm-up, before paired use)."""
        with torch.cuda.device(self.dev):
            _C.check(_C.lib.nafae_gate_sync(_C.ptr(self.gate), _C.stream(self.dev)), "nafae_gate_sync")

    def wait_gate(self, slot):
        """Hold the current stream until this step's gated RoIAlign kernel owns its SMs."""
        with torch.cuda.device(self.dev):
            _C.check(_C.lib.nafae_gate_wait(_C.ptr(self.gate), int(slot), _C.stream(self.dev)),
                     "nafae_gate_wait")

    def run_detector(self):
        """Detector-side half: proposal tail -> RoIAlignAvg.  Frozen in NAFAE (model.py:651,673,
        706-707): independent of the trainable weights, so it may run ahead of / concurrently with
        earlier batches' head (see `capture_pipelined`)."""
        self.run_tail()
        self.run_align(gated=False)  # nobody waits on the gate in the sequential form

    def run_head(self, backward=None):
        """Head half: similarity + losses forward, then backward to dL/dvis_feats, dL/dword_feats.
        (bridge: fc6/fc7 + VisEbd / WordEbd run between the halves in the caller's PyTorch code and
        produce vis_feats / word_feats; they are outside this path)"""
        Na, Ns, Nb, Ne, D = self.dims
        L, P = _C.lib, _C.ptr
        s = _C.stream(self.dev)
        do_bwd = self.train if backward is None else backward
        with torch.cuda.device(self.dev):
            fwd = L.nafae_ground_forward_tc if self.tensor_cores else L.nafae_ground_forward
            _C.check(fwd(P(self.vis_feats), P(self.word_feats), P(self.lens),
                                            Na, Ns, Nb, Ne, D, self.Delta, self.vis_lam,
                                            int(self.train), P(self.D_ind), P(self.D_sim),
                                            P(self.loss), P(self.ws), self.ws.numel() * 4, s),
                     "nafae_ground_forward")
            if do_bwd:
                _C.check(L.nafae_ground_backward(P(self.grad_loss), P(self.vis_feats),
                                                 P(self.word_feats), P(self.lens), Na, Ns, Nb, Ne,
                                                 D, self.Delta, self.vis_lam, int(self.train),
                                                 P(self.D_ind), P(self.D_sim), P(self.grad_vis),
                                                 P(self.grad_word), P(self.ws),
                                                 self.ws.numel() * 4, s), "nafae_ground_backward")

    def run(self, backward=None):
        """Enqueue the whole step on the current stream.  No host sync, no allocation."""
        self.run_detector()
        self.run_head(backward)

    def kernels_per_step(self):
        # the tcgen05 scoring path is two kernels (ground_p1_tc_kernel, ground_p23_kernel) instead of ground_fwd
        return (self.KERNELS_PER_STEP_TRAIN if self.train else self.KERNELS_PER_STEP_EVAL) + int(self.tensor_cores)

    def capture(self):
        """Capture run() into a CUDA graph (launch-bound at this size: 5 kernels, ~100 us)."""
        self.run()  # warm up (sets func attributes, loads modules) outside capture
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.run()
        self.graph = g
        return g

    def replay(self):
        self.graph.replay()


def capture_pipelined(align_step, next_step, streams, extra_branch=None, gate_head=False):
    """CUDA graph of one software-pipelined training step (three concurrent branches):

        A (capturing stream): RoIAlign of `align_step`        -- batch k+1, rois from the last replay
        B (streams[0])      : proposal tail of `next_step`    -- batch k+2, rois for the next replay
        C (streams[1])      : head (DVSA fwd + bwd) of `next_step` -- batch k
        D (optional)        : `extra_branch(cur)` on its own stream (gradient all-reduce)

    The detector is frozen in NAFAE, so a later batch's NMS / RoIAlign do not depend on an earlier
    batch's weight update: the stages of consecutive batches overlap like a data-loader prefetch.
    Every replay executes exactly one proposal tail, one RoIAlign and one head; dependencies inside
    a batch (tail -> RoIAlign through `rois`, fwd -> bwd) are preserved because graphs serialise
    and the two buffer sets alternate.  The slab kernel is persistent: leave a few SMs to the head
    kernels with `nafae_set_reserved_sms` (the tail's small CTAs co-reside with it).

    All branches start at the same instant and the block scheduler places CTAs breadth-first, so
    the head's first kernel lands on every SM and the 210 KB persistent RoIAlign CTAs start late
    on the SMs where its long-lived CTAs sit.  A branch that must NOT spread over the GPU (the
    all-reduce: its CTAs spin on peers) waits behind the kernel's residency gate (`wait_gate`);
    `gate_head=True` does the same for the head (measured slower: its kernels want the whole GPU)."""
    align_step.sync_gate()  # eager: earlier unpaired gated launches must not satisfy this graph's waits
    torch.cuda.current_stream().synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        capture_pipelined_body(align_step, next_step, streams, extra_branch, gate_head)
    return g


def capture_pipelined_body(align_step, next_step, streams, extra_branch=None, gate_head=False):
    """The fork / join of one pipelined step on the capturing stream (call inside torch.cuda.graph;
    several bodies in a row make a multi-step graph that amortises the graph-launch latency)."""
    cur = torch.cuda.current_stream()
    joined = []

    def head():
        if gate_head:
            align_step.wait_gate(0)
        next_step.run_head()
    for st, fn in ((streams[0], next_step.run_tail), (streams[1], head)):
        st.wait_stream(cur)
        with torch.cuda.stream(st):
            fn()
        joined.append(st)
    if extra_branch is not None:
        extra_stream = extra_branch(cur)  # a gated extra branch calls align_step.wait_gate(1) itself
        if extra_stream is not None:
            joined.append(extra_stream)
    align_step.run_align()
    for st in joined:
        cur.wait_stream(st)


def capture_pipelined_exact(steps, num_steps, streams, allreduce=None, comm=None):
    """CUDA graph of `num_steps` (even) consecutive software-pipelined steps that carries only the TRUE
    data dependencies between them, instead of a full join after every step (`capture_pipelined`):

        step s, j = s & 1:   RoIAlign of set j  ||  proposal tail of set 1-j  ||  head of set 1-j
                             [ ||  all-reduce of bucket j ]

        RoIAlign(set j)   @ s  after  tail(set j) @ s-1                 (reads its rois)
        tail(set 1-j)     @ s  after  RoIAlign(set 1-j) @ s-1           (rewrites the rois that kernel read)
        head(set 1-j)     @ s  after  head(set 1-j) @ s-2               (same buffers; stream order)
                               after  all-reduce(bucket 1-j) @ s-1      (rewrites the bucket)
        all-reduce(j)     @ s  after  head(set j) @ s-1                 (reduces what it wrote)

    RoIAlign kernels follow each other on the capturing stream, tails on streams[0], the heads of the two
    sets on streams[1] / streams[2], all-reduces on `comm` (make it a high-priority stream).  The chains
    slide against each other inside the graph -- the heads run ahead, the RoIAlign kernels run back to
    back -- so a step costs what the slowest CHAIN needs per step, not the slowest chain plus a join and a
    replay gap.  Every replay still executes exactly `num_steps` of each kernel.  `allreduce` = the two
    bucket objects (bucket j = gradients of set j), launched ungated: with several aligned RoIAlign launches
    in flight a residency-gate wait could pair with none of them."""
    if num_steps < 2 or num_steps % 2:
        raise ValueError("num_steps must be even and >= 2")
    if len(streams) < 3:
        raise ValueError("capture_pipelined_exact needs three side streams (tails, heads of set 0 / set 1)")
    if allreduce is not None and comm is None:
        raise ValueError("an all-reduce branch needs its own stream")
    tails, heads = streams[0], (streams[1], streams[2])
    side = [tails, heads[0], heads[1]] + ([comm] if allreduce is not None else [])
    torch.cuda.current_stream().synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        cur = torch.cuda.current_stream()
        for st in side:
            st.wait_stream(cur)
        ev_align = ev_tail = None          # of the previous step
        ev_head, ev_ar = [None, None], [None, None]   # latest head / all-reduce touching set / bucket i

        def mark(stream):
            ev = torch.cuda.Event()
            ev.record(stream)
            return ev
        for s in range(num_steps):
            j = s & 1
            if ev_align is not None:
                tails.wait_event(ev_align)
            with torch.cuda.stream(tails):
                steps[1 - j].run_tail()
                new_tail = mark(tails)
            if ev_tail is not None:
                cur.wait_event(ev_tail)
            steps[j].run_align(gated=False)
            new_align = mark(cur)
            if allreduce is not None:
                if ev_head[j] is not None:
                    comm.wait_event(ev_head[j])
                with torch.cuda.stream(comm):
                    allreduce[j].launch()
                    ev_ar[j] = mark(comm)
            h = heads[1 - j]
            if ev_ar[1 - j] is not None:
                h.wait_event(ev_ar[1 - j])
            with torch.cuda.stream(h):
                steps[1 - j].run_head()
                ev_head[1 - j] = mark(h)
            ev_align, ev_tail = new_align, new_tail
        for st in side:
            cur.wait_stream(st)
    return g
