"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch).

Video segments shard across ranks with no forward exchange; the only collective
of a training step is one all-reduce (SUM, then /world) over a flat fp32 bucket of the trainable
gradients -- vis_ebd.fc1 (512x4096+512), word_ebd.fc1 (512xglove+512), word_ebd.bn (2x512): 2.20 M
floats = 8.8 MB at glove_dim 200 (reference model.py:616-642, optimiser at model.py:1030-1036).
The all-reduce is issued on a side stream so it overlaps the next step's NMS / RoIAlign, which do
not depend on gradients.
"""
import os

import torch
import torch.distributed as dist


# SMs left to the collective while the slab kernel runs -- NCCL and the 256-thread peer-memory
# all-reduce only; the default 128-thread variant co-resides with the slab CTAs and needs none
COMM_SMS = int(os.environ.get("NAFAE_COMM_SMS", "16"))


# largest world size for which make_allreduce("auto") prefers the peer-memory kernel over NVLS
AUTO_PEER_MAX_WORLD = int(os.environ.get("NAFAE_AUTO_PEER_MAX_WORLD", "4"))


def trainable_grad_elems(vis_fc_dim=4096, glove_dim=200, ebd_dim=512):
    return (ebd_dim * vis_fc_dim + ebd_dim) + (ebd_dim * glove_dim + ebd_dim) + 2 * ebd_dim


def init_from_env(backend=None):
    """Join the process group torchrun describes; returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
            # the all-reduce overlaps the persistent RoIAlign kernel: bound the SMs NCCL takes and
            # keep that many free (see nafae_set_reserved_sms in include/nafae_b200.h)
            os.environ.setdefault("NCCL_MAX_NCHANNELS", str(COMM_SMS))
            from . import _C
            _C.lib.nafae_set_reserved_sms(COMM_SMS)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def shard_segments(num_segments, rank, world):
    """Contiguous shard [begin, end) of the segment list for `rank` (frames of a segment are never
    split across ranks)."""
    base, rem = divmod(num_segments, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_dets(dets):
    """Inference sweep: every rank grounds its own contiguous shard of the
    segment list (`shard_segments`) and records detections with GLOBAL image ids; this merges the
    four parallel lists `[img_ids, labels, boxes, confs]` of all ranks in rank order -- the same lists
    a single process walking the segments in order would have produced (model.py:972), ready for
    `evaluate.evaluate_box` / `evaluate.save_dets`.  One `all_gather_object` (any backend); returns the
    merged lists on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [list(x) for x in dets]
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, [list(x) for x in dets])
    return [[item for part in parts for item in part[k]] for k in range(4)]


class GradBucket(object):
    """Flat fp32 gradient bucket with an overlapped all-reduce.

    ``views(shapes)`` hands out non-overlapping views for the parameter gradients; ``allreduce_async``
    launches SUM/world on a side stream after the producing stream's current work; ``wait`` makes the
    consuming stream wait for it (no host sync on either side).
    """

    def __init__(self, numel, device, world=None):
        self.buf = torch.zeros((numel,), dtype=torch.float32, device=device)
        self.world = world if world is not None else (dist.get_world_size()
                                                      if dist.is_initialized() else 1)
        self.stream = torch.cuda.Stream(device) if self.buf.is_cuda else None
        self._done = None

    def views(self, shapes):
        out, off = [], 0
        for shp in shapes:
            n = 1
            for s in shp:
                n *= int(s)
            out.append(self.buf[off:off + n].view(*shp))
            off += n
        if off > self.buf.numel():
            raise ValueError("bucket too small")
        return out

    def allreduce_async(self):
        if self.world <= 1:
            return
        if self.stream is None:  # CPU / gloo (tests; gloo has no AVG)
            dist.all_reduce(self.buf, op=dist.ReduceOp.SUM)
            self.buf.div_(self.world)
            return
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            dist.all_reduce(self.buf, op=dist.ReduceOp.AVG)  # SUM/world inside NCCL: one kernel
            self._done = torch.cuda.Event()
            self._done.record()

    def wait(self):
        if self._done is not None:
            torch.cuda.current_stream().wait_event(self._done)
            self._done = None


class _RawCudaArray(object):
    """Minimal __cuda_array_interface__ carrier so torch can view memory the C library owns."""

    def __init__(self, ptr, numel):
        self.__cuda_array_interface__ = dict(shape=(int(numel),), typestr="<f4", data=(int(ptr), False),
                                             version=2)


class PeerAllReduce(object):
    """Flat fp32 gradient bucket in a symmetric (CUDA-IPC mapped) buffer + the two-shot NVLink
    all-reduce kernel of csrc/allreduce.cu.  One process per GPU, single node, world <= 8.

    `buf` is a torch view of this rank's bucket (write gradients into `views(...)` of it);
    `launch()` enqueues ONE kernel on the current stream (graph-capturable, no NCCL involved);
    every rank must call it the same number of times.  torch.distributed is only used once, to
    exchange the 64-byte IPC handles."""

    kind = "peer"

    def __init__(self, numel, device, rank=None, world=None, num_ctas=None, cta_threads=None,
                 variant=None, width=0):
        import ctypes
        from . import _C
        self._C, self._ct = _C, ctypes
        # include/nafae_b200.h: NAFAE_AR_VARIANT(v) | NAFAE_AR_WIDTH(w)
        # variant 1 (3-slot ring, larger chunks, one issuing lane per peer) measured faster than variant
        # 0 at every width (world 2: 34.9 vs 38.7 us; <W=8>: 52.5 vs 90.1 us for the 8.8 MB bucket)
        if variant is None:
            variant = int(os.environ.get("NAFAE_AR_VARIANT", "1"))
        self.flags = (int(variant) & 0xf) | ((int(width) & 0xff) << 8)
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.dev = torch.device(device)
        pad = 4 * self.world
        self.count = (int(numel) + pad - 1) // pad * pad
        self.numel = int(numel)
        # cta_threads 0 = bulk-copy (TMA) kernel, one CTA per SM left free by nafae_set_reserved_sms;
        # 256 / 128 = per-thread-load kernels; see csrc/allreduce.cu
        self.cta_threads = int(cta_threads if cta_threads is not None else
                               os.environ.get("NAFAE_AR_THREADS", "0"))
        self.num_ctas = int(num_ctas if num_ctas is not None else
                            os.environ.get("NAFAE_AR_CTAS", {0: str(COMM_SMS), 128: "128"}.get(self.cta_threads, "96")))
        nbytes = int(_C.lib.nafae_ar_buffer_bytes(self.count, self.world))
        own = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        with torch.cuda.device(self.dev):
            _C.check(_C.lib.nafae_ar_alloc(nbytes, ctypes.byref(own), handle), "nafae_ar_alloc")
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle))
            self._ptrs = (ctypes.c_void_p * self.world)()
            self._opened = []
            for r in range(self.world):
                if r == self.rank:
                    self._ptrs[r] = own.value
                else:
                    peer = ctypes.c_void_p()
                    hb = (ctypes.c_ubyte * 64).from_buffer_copy(handles[r])
                    _C.check(_C.lib.nafae_ar_open(hb, ctypes.byref(peer)), "nafae_ar_open")
                    self._ptrs[r] = peer.value
                    self._opened.append(peer.value)
        self._own = own.value
        off = int(_C.lib.nafae_ar_data_offset())
        self.buf = torch.as_tensor(_RawCudaArray(self._own + off, self.count), device=self.dev)[: self.numel]
        dist.barrier()

    def views(self, shapes):
        out, off = [], 0
        for shp in shapes:
            n = 1
            for s in shp:
                n *= int(s)
            out.append(self.buf[off:off + n].view(*shp))
            off += n
        if off > self.buf.numel():
            raise ValueError("bucket too small")
        return out

    def launch(self):
        """All-reduce (AVG) the bucket in place on the current stream."""
        if self.world <= 1:
            return
        with torch.cuda.device(self.dev):
            st = self._C.lib.nafae_allreduce_avg(self._ptrs, self.rank, self.world, self.count,
                                                 self.num_ctas, self.cta_threads, self.flags,
                                                 self._C.stream(self.dev))
        self._C.check(st, "nafae_allreduce_avg")

    def close(self):
        torch.cuda.synchronize(self.dev)
        dist.barrier()
        for p in self._opened:
            self._C.lib.nafae_ar_close(self._ct.c_void_p(p))
        self._opened = []
        dist.barrier()
        if self._own:
            self.buf = None
            self._C.lib.nafae_ar_free(self._ct.c_void_p(self._own))
            self._own = None


def _share_fd_from_root(fd, rank, world, root=0):
    """Hand one file descriptor from `root` to every other rank of the (single-node) group over
    an abstract AF_UNIX socket with SCM_RIGHTS.  Returns the received descriptor (root: `fd`)."""
    import socket
    import time
    name = [None]
    srv = None
    if rank == root:
        srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        name[0] = "\0nafae-b200-mc-%d-%d" % (os.getpid(), int(time.time() * 1e6) & 0xffffffff)
        srv.bind(name[0])
        srv.listen(world)
    dist.broadcast_object_list(name, src=root)
    got = fd
    if rank == root:
        for _ in range(world - 1):
            conn, _addr = srv.accept()
            socket.send_fds(conn, [b"f"], [fd])
            conn.recv(1)  # the peer has the descriptor: safe to close
            conn.close()
        srv.close()
    else:
        c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        for attempt in range(200):
            try:
                c.connect(name[0])
                break
            except (ConnectionRefusedError, FileNotFoundError):
                time.sleep(0.01)
        else:
            raise RuntimeError("could not reach the multicast root's socket")
        _msg, fds, _flags, _addr = socket.recv_fds(c, 16, 1)
        c.send(b"k")
        c.close()
        if not fds:
            raise RuntimeError("no file descriptor received from the multicast root")
        got = fds[0]
    return got


def multicast_supported(device=None):
    """True when the current device can bind memory to an NVSwitch multicast object (NVLS)."""
    from . import _C
    if not torch.cuda.is_available():
        return False
    with torch.cuda.device(device if device is not None else torch.cuda.current_device()):
        return int(_C.lib.nafae_mc_supported()) == 1


class MulticastAllReduce(object):
    """Flat fp32 gradient bucket in NVSwitch MULTICAST memory + the in-switch (NVLS) all-reduce
    kernel of csrc/allreduce.cu (`allreduce_mc_kernel`: multimem.ld_reduce / multimem.st).

    Same interface as `PeerAllReduce`.  torch.distributed is only used during construction
    (to pass the multicast object's file descriptor around and for two host barriers)."""
    kind = "multicast"

    def __init__(self, numel, device, rank=None, world=None, num_ctas=None, cta_threads=None):
        import ctypes
        from . import _C
        self._C, self._ct = _C, ctypes
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.dev = torch.device(device)
        pad = 4 * self.world
        self.count = (int(numel) + pad - 1) // pad * pad
        self.numel = int(numel)
        self.cta_threads = int(cta_threads if cta_threads is not None else
                               os.environ.get("NAFAE_MC_THREADS", "512"))
        self.num_ctas = int(num_ctas if num_ctas is not None else os.environ.get("NAFAE_MC_CTAS", "8"))
        nbytes = int(_C.lib.nafae_mc_buffer_bytes(self.count, self.world))
        h = ctypes.c_void_p()
        self._h = None
        with torch.cuda.device(self.dev):
            # every rank reports whether it can go on BEFORE anyone blocks in a collective step
            ok = [None] * self.world
            fd = ctypes.c_int(-1)
            err = ""
            if self.rank == 0:
                st = int(_C.lib.nafae_mc_create(self.world, nbytes, ctypes.byref(h), ctypes.byref(fd)))
                if st != 1:
                    err = _C.last_error()
            dist.all_gather_object(ok, err)
            if ok[0]:
                raise RuntimeError("nafae_mc_create failed on rank 0: %s" % ok[0])
            got = _share_fd_from_root(int(fd.value), self.rank, self.world)
            err = ""
            if self.rank != 0:
                st = int(_C.lib.nafae_mc_import(int(got), self.world, nbytes, ctypes.byref(h)))
                if st != 1:
                    err = _C.last_error()
            try:
                os.close(int(got))
            except OSError:
                pass
            if not err:
                if int(_C.lib.nafae_mc_add_device(h)) != 1:
                    err = _C.last_error()
            dist.all_gather_object(ok, err)
            if any(ok):
                raise RuntimeError("multicast setup failed: %s" % [e for e in ok if e])
            uc, mc = ctypes.c_void_p(), ctypes.c_void_p()
            err = ""
            if int(_C.lib.nafae_mc_bind(h, ctypes.byref(uc), ctypes.byref(mc))) != 1:
                err = _C.last_error()
            dist.all_gather_object(ok, err)
            if any(ok):
                raise RuntimeError("multicast bind failed: %s" % [e for e in ok if e])
        self._h, self._uc, self._mc = h, uc.value, mc.value
        off = int(_C.lib.nafae_ar_data_offset())
        self.buf = torch.as_tensor(_RawCudaArray(self._uc + off, self.count), device=self.dev)[: self.numel]
        dist.barrier()

    views = PeerAllReduce.views

    def launch(self):
        """All-reduce (AVG) the bucket in place on the current stream."""
        if self.world <= 1:
            return
        with torch.cuda.device(self.dev):
            st = self._C.lib.nafae_allreduce_mc(self._ct.c_void_p(self._uc), self._ct.c_void_p(self._mc),
                                                self.rank, self.world, self.count, self.num_ctas,
                                                self.cta_threads, self._C.stream(self.dev))
        self._C.check(st, "nafae_allreduce_mc")

    def timed_out(self):
        """Host-synchronising: did a cross-GPU wait of this bucket ever time out?"""
        with torch.cuda.device(self.dev):
            return int(self._C.lib.nafae_allreduce_mc_error(self._ct.c_void_p(self._uc))) != 0

    def close(self):
        torch.cuda.synchronize(self.dev)
        dist.barrier()
        if self._h is not None:
            self.buf = None
            with torch.cuda.device(self.dev):
                self._C.lib.nafae_mc_free(self._h)
            self._h = None
        dist.barrier()


def make_allreduce(numel, device, kind="auto", peer_kw=None, mc_kw=None):
    """The gradient all-reduce of the data-parallel step: `multicast` (NVLS, in-switch reduction)
    when every rank's device supports it, else `peer` (bulk-copy two-shot over CUDA-IPC peer
    memory).  `kind` = "auto" | "multicast" | "peer"; all ranks take the same decision."""
    if kind not in ("auto", "multicast", "peer"):
        raise ValueError("kind must be auto, multicast or peer")
    if kind == "auto" and dist.get_world_size() <= AUTO_PEER_MAX_WORLD:
        # measured (profiles/RESULTS.md, round 2): in the pipelined step the bulk-copy peer kernel is
        # ahead at 2 and 4 ranks (59.0 vs 61.5 and 60.4 vs 61.5 us / step) and level at 8 (59.9 vs 60.1),
        # where the NVLS kernel needs half the SMs (8 CTAs) and the RoIAlign kernel keeps 122 instead of 112
        kind = "peer"
    if kind != "peer":
        mine = multicast_supported(device)
        flags = [None] * dist.get_world_size()
        dist.all_gather_object(flags, bool(mine))
        if all(flags):
            try:
                return MulticastAllReduce(numel, device, **(mc_kw or {}))
            except RuntimeError as e:  # raised collectively (every rank sees the same failure)
                if kind == "multicast":
                    raise
                import sys
                print("nafae_b200: NVLS multicast setup failed (%s); using the peer-memory all-reduce" % e,
                      file=sys.stderr)
        elif kind == "multicast":
            raise RuntimeError("NVSwitch multicast is not supported on every rank: %s" % flags)
    return PeerAllReduce(numel, device, **(peer_kw or {}))


def capture_step_with_allreduce(step, reduce_bucket, side_stream):
    """One CUDA graph per training step: the whole hot path of `step` on the capturing stream and,
    as a parallel branch, the all-reduce (AVG) of `reduce_bucket` -- the gradients the PREVIOUS step
    produced.  Replaying graph i+1 therefore overlaps step i's gradient all-reduce with step i+1's
    NMS / RoIAlign inside a single launch; consecutive replays serialise, so a bucket is never
    rewritten before its all-reduce has finished."""
    step.run()  # warm up outside capture
    if reduce_bucket is not None and reduce_bucket.world > 1:
        dist.all_reduce(reduce_bucket.buf, op=dist.ReduceOp.AVG)  # communicator warm-up
    torch.cuda.synchronize(step.dev)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        cur = torch.cuda.current_stream()
        if reduce_bucket is not None and reduce_bucket.world > 1:
            side_stream.wait_stream(cur)
            with torch.cuda.stream(side_stream):
                dist.all_reduce(reduce_bucket.buf, op=dist.ReduceOp.AVG)
        step.run()
        if reduce_bucket is not None and reduce_bucket.world > 1:
            cur.wait_stream(side_stream)
    step.graph = g
    return g
