"""Multi-rank worker of tests/test_multi_gpu.py (one process per GPU, launched by torch.distributed.run).

Checks, on every rank:
  1. every all-reduce variant of csrc/allreduce.cu -- the bulk-copy peer-memory kernel at every
     compile-time width and both ring variants, the per-thread-load kernels, the NVLS multimem
     kernel when the box supports multicast -- against an fp64 host mean of all ranks' inputs,
     eagerly (3 launches in a row) and as CUDA-graph replays, with bit-identical replicas;
  2. three software-pipelined data-parallel steps (bench.py's schedule): the averaged dL/dword
     equals the mean of the per-rank oracle gradients and is identical on every replica;
  3. three HeadTrainer steps (backward -> all-reduce -> clip -> Adam): parameters stay bit-identical
     across replicas and change.
With NAFAE_MGPU_TIME=1 it also prints the standalone time of every variant (graph replays).
Prints MGPU_OK on rank 0 when everything passed.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from nafae_b200 import parallel, synth  # noqa: E402

rank, world, local = parallel.init_from_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
TIME = os.environ.get("NAFAE_MGPU_TIME") == "1"
ONLY_EXACT = os.environ.get("NAFAE_MGPU_ONLY") == "exact"  # dev: just the dependency-exact DP graphs
N = parallel.trainable_grad_elems()
failures = []


def log(msg):
    if rank == 0:
        print(msg, flush=True)


def gather_all(t):
    """(world, n) copy of `t` from every rank (NCCL; the checker, not the thing under test)."""
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t.contiguous())
    return torch.stack(out)


def check_allreduce(name, ar):
    g = torch.Generator(device="cpu").manual_seed(1000 + 17 * rank)
    ok = True
    for it in range(3):  # eager, back to back: exercises the epoch handling
        x = torch.randn(ar.numel, generator=g).to(dev) * (1.0 + it)
        ar.buf.copy_(x)
        want = gather_all(x).double().mean(0)
        torch.cuda.synchronize()
        dist.barrier()
        ar.launch()
        torch.cuda.synchronize()
        got = ar.buf.clone()
        err = ((got.double() - want).abs().max() / want.abs().max()).item()
        allr = gather_all(got)
        same = bool((allr.view(torch.int32) == allr[0].view(torch.int32)).all().item())
        if not (err < 1e-6 and same):
            ok = False
            failures.append("%s: iter %d rel err %.3e, replicas identical %s" % (name, it, err, same))
    # CUDA-graph replays (how the step uses it)
    gr = torch.cuda.CUDAGraph()
    x = torch.randn(ar.numel, generator=g).to(dev)
    ar.buf.copy_(x)
    torch.cuda.synchronize()
    dist.barrier()
    with torch.cuda.graph(gr):
        ar.launch()
    want = gather_all(x).double().mean(0)
    for k in range(4):  # AVG of identical replicas is a fixed point after the first replay
        gr.replay()
    torch.cuda.synchronize()
    err = ((ar.buf.double() - want).abs().max() / want.abs().max()).item()
    if not err < 1e-6:
        ok = False
        failures.append("%s: graph replay rel err %.3e" % (name, err))
    us = None
    if TIME:
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(20):
            gr.replay()
        torch.cuda.synchronize()
        dist.barrier()
        e0.record()
        for _ in range(200):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 200 * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        us = float(t.item())
    if hasattr(ar, "timed_out") and ar.timed_out():
        ok = False
        failures.append("%s: a cross-GPU wait timed out" % name)
    log("allreduce %-34s world %d  %s%s" % (name, world, "ok" if ok else "FAILED",
                                            "  %.1f us" % us if us is not None else ""))
    ar.close()


# ------------------------------------------------------------------ 1. all-reduce variants ----
widths = [w for w in (2, 4, 8) if w >= world] if not ONLY_EXACT else []
for w in widths:
    for v in (0, 1):
        check_allreduce("peer bulk-copy <W=%d, V=%d> x16" % (w, v),
                        parallel.PeerAllReduce(N, dev, num_ctas=16, cta_threads=0, variant=v, width=w))
if not ONLY_EXACT:
    check_allreduce("peer bulk-copy auto width x8", parallel.PeerAllReduce(N, dev, num_ctas=8, cta_threads=0))
    check_allreduce("peer per-thread 256 x96", parallel.PeerAllReduce(N, dev, num_ctas=96, cta_threads=256))
    check_allreduce("peer per-thread 128 x128", parallel.PeerAllReduce(N, dev, num_ctas=128, cta_threads=128))
mc_flags = [None] * world
dist.all_gather_object(mc_flags, parallel.multicast_supported(dev))
HAVE_MC = all(mc_flags)
log("NVSwitch multicast supported on every rank: %s" % HAVE_MC)
if HAVE_MC and not ONLY_EXACT:
    for ctas, thr in ((16, 512), (8, 512), (4, 1024), (32, 256), (2, 512)):
        check_allreduce("multicast (NVLS) %dx%d" % (ctas, thr),
                        parallel.MulticastAllReduce(N, dev, num_ctas=ctas, cta_threads=thr))
    ar = parallel.make_allreduce(N, dev)
    want_kind = "peer" if world <= parallel.AUTO_PEER_MAX_WORLD else "multicast"
    if ar.kind != want_kind:
        failures.append("make_allreduce(auto) picked %s, expected %s at world %d" % (ar.kind, want_kind, world))
    ar.close()
elif not HAVE_MC:
    ar = parallel.make_allreduce(N, dev)
    if ar.kind != "peer":
        failures.append("make_allreduce(auto) must fall back to the peer-memory kernel")
    ar.close()

# -------------------------------------------------- 2. pipelined data-parallel steps ----
from nafae_b200 import _C  # noqa: E402
from nafae_b200.pipeline import GroundingStep, capture_pipelined  # noqa: E402
from oracle import dvsa as odvsa  # noqa: E402  (checker)

c = dict(synth.CONFIGS["cfg2"])
c.update(C=32, n=600)  # small maps / fewer proposals: the DVSA half is the full cfg2 shape
host = [synth.make_batch(c, 500 + 10 * rank + i) for i in range(2)]
want = []
for hb in host:
    ref = odvsa.dvsa_forward_backward(hb["vis_feats"], hb["word_feats"], hb["lens"], c["Na"], c["Nb"],
                                      c["Ne"], c["Delta"], c["vis_lam"], "train")
    mine = torch.from_numpy(np.ascontiguousarray(ref["grad_word"])).to(dev)
    want.append(gather_all(mine).double().mean(0))


def run_dp(kind, tensor_cores, gated):
    """bench.py's schedule: head variant (fp32 FMA / tcgen05), all-reduce behind the residency gate or not,
    the all-reduce branch on a high-priority stream."""
    steps = [GroundingStep(c["Na"], c["Ns"], c["Nb"], c["Ne"], c["D"], c["C"], c["H"], c["W"], c["n"],
                           pre_nms_topn=c["pre"], Delta=c["Delta"], vis_lam=c["vis_lam"], train=True,
                           device=dev, tensor_cores=tensor_cores) for _ in range(2)]
    buckets = [parallel.make_allreduce(N, dev, kind=kind) for _ in range(2)]
    for st, b, hb in zip(steps, buckets, host):
        st.grad_word = b.views([(st.NQ, c["D"])])[0]
        st.load(hb)
        st.run()
    prev = _C.lib.nafae_set_reserved_sms(32)
    side = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
    comm = torch.cuda.Stream(dev, priority=-1)
    torch.cuda.synchronize()
    graphs = []
    for j in range(2):
        def ar_branch(cur, j=j):
            comm.wait_stream(cur)
            with torch.cuda.stream(comm):
                if gated:
                    steps[j].wait_gate(1)
                buckets[j].launch()
            return comm
        graphs.append(capture_pipelined(steps[j], steps[1 - j], side, ar_branch))
    torch.cuda.synchronize()
    dist.barrier()
    bad = 0
    for k in range(6):
        graphs[k & 1].replay()
        # replay k: head of set 1-k&1 wrote bucket[1-k&1]; the all-reduce averaged bucket[k&1] (filled by
        # the previous replay's head, or by the warm-up run for k = 0)
        torch.cuda.synchronize()
        j = k & 1
        got = steps[j].grad_word.double()
        err = ((got - want[j]).abs().max() / want[j].abs().max()).item()
        allr = gather_all(steps[j].grad_word.reshape(-1))
        same = bool((allr.view(torch.int32) == allr[0].view(torch.int32)).all().item())
        if not (err < 1e-4 and same):
            bad += 1
            failures.append("pipelined DP step %d (%s, tcgen05 head %s, gated %s): grad_word rel err %.3e, replicas "
                            "identical %s" % (k, buckets[0].kind, tensor_cores, gated, err, same))
    log("pipelined DP steps (%s all-reduce, %s, %s head): %s" % (
        buckets[0].kind, "gated" if gated else "ungated", "tcgen05" if tensor_cores else "fma",
        "ok" if not bad else "FAILED"))
    _C.lib.nafae_set_reserved_sms(prev)
    for b in buckets:
        b.close()
    del graphs


def run_dp_exact(kind, tensor_cores):
    """bench.py's default schedule: several steps per graph with only the true dependencies between them,
    the all-reduce ungated on a high-priority stream (pipeline.capture_pipelined_exact)."""
    from nafae_b200.pipeline import capture_pipelined_exact
    steps = [GroundingStep(c["Na"], c["Ns"], c["Nb"], c["Ne"], c["D"], c["C"], c["H"], c["W"], c["n"],
                           pre_nms_topn=c["pre"], Delta=c["Delta"], vis_lam=c["vis_lam"], train=True,
                           device=dev, tensor_cores=tensor_cores) for _ in range(2)]
    buckets = [parallel.make_allreduce(N, dev, kind=kind) for _ in range(2)]
    for st, b, hb in zip(steps, buckets, host):
        st.grad_word = b.views([(st.NQ, c["D"])])[0]
        st.load(hb)
        st.run()
    prev = _C.lib.nafae_set_reserved_sms(32)
    side = [torch.cuda.Stream(dev) for _ in range(3)]
    comm = torch.cuda.Stream(dev, priority=-1)
    torch.cuda.synchronize()
    dist.barrier()
    g = capture_pipelined_exact(steps, 4, side, allreduce=buckets, comm=comm)
    torch.cuda.synchronize()
    dist.barrier()
    bad = 0
    for k in range(3):
        g.replay()
    torch.cuda.synchronize()
    buckets[0].launch()  # flush: the last step's head wrote bucket 0, nobody has reduced it yet
    torch.cuda.synchronize()
    for j in range(2):
        got = steps[j].grad_word.double()
        err = ((got - want[j]).abs().max() / want[j].abs().max()).item()
        allr = gather_all(steps[j].grad_word.reshape(-1))
        same = bool((allr.view(torch.int32) == allr[0].view(torch.int32)).all().item())
        if not (err < 1e-4 and same):
            bad += 1
            failures.append("dependency-exact DP graph (%s, tcgen05 head %s): bucket %d rel err %.3e, replicas "
                            "identical %s" % (buckets[0].kind, tensor_cores, j, err, same))
    log("dependency-exact 4-step DP graph (%s all-reduce, %s head): %s" % (
        buckets[0].kind, "tcgen05" if tensor_cores else "fma", "ok" if not bad else "FAILED"))
    _C.lib.nafae_set_reserved_sms(prev)
    for b in buckets:
        b.close()
    del g


if not ONLY_EXACT:
    run_dp("peer", False, True)
    run_dp("peer", True, True)
run_dp_exact("peer", True)
if HAVE_MC:
    if not ONLY_EXACT:
        run_dp("multicast", False, True)
        run_dp("multicast", True, False)
    run_dp_exact("multicast", True)

# ------------------------------------------------------------- 3. HeadTrainer steps ----
from nafae_b200.bridge import VisEbd, WordEbd  # noqa: E402
from nafae_b200.train_step import HeadTrainer  # noqa: E402

args = types.SimpleNamespace(vis_fc_dim=4096, glove_dim=200, word_ebd_dim=512, dropout_rate=0.0)
torch.manual_seed(7)  # same initial weights on every replica
vis_ebd, word_ebd = VisEbd(args).to(dev), WordEbd(args).to(dev)
n_par = HeadTrainer.numel(vis_ebd, word_ebd)
ar = parallel.make_allreduce(n_par, dev)
tr = HeadTrainer(vis_ebd, word_ebd, Na=8, Nb=20, Ne=13, Delta=10.0, vis_lam=4.13, allreduce=ar)
g = torch.Generator(device="cpu").manual_seed(900 + rank)  # different data per rank
p0 = tr.flat_param.clone()
for it in range(3):
    fc = (torch.randn(8 * 5 * 20, 4096, generator=g) * 30).to(dev)
    gl = (torch.randn(8 * 13, 200, generator=g) * 0.4).to(dev)
    lens = [int(x) for x in torch.randint(1, 6, (8,), generator=g)]
    tr.step(fc, gl, lens)
    torch.cuda.synchronize()
    allp = gather_all(tr.flat_param)
    same = bool((allp.view(torch.int32) == allp[0].view(torch.int32)).all().item())
    if not same:
        failures.append("HeadTrainer step %d: replicas diverged" % it)
if not (tr.flat_param - p0).abs().max().item() > 0:
    failures.append("HeadTrainer: parameters did not change")
if not torch.isfinite(tr.flat_param).all().item():
    failures.append("HeadTrainer: non-finite parameters")
log("HeadTrainer 3 DP steps (%s all-reduce): replicas identical: %s" % (ar.kind, not any("HeadTrainer" in f for f in failures)))
ar.close()

bad = [None] * world
dist.all_gather_object(bad, failures)
allbad = [f for fs in bad for f in fs]
if rank == 0:
    for f in allbad:
        print("FAIL: " + f, flush=True)
    if not allbad:
        print("MGPU_OK", flush=True)
dist.destroy_process_group()
sys.exit(1 if allbad else 0)
