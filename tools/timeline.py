"""Dev tool: CTA-level timeline of the software-pipelined step (needs libnafae_b200_trace.so:
make -C nafae_b200/csrc trace).  Prints, for the last replay, when each kernel's CTAs started and
ended relative to the first CTA of the replay, and the RoIAlign kernel's own / stolen unit counts."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["NAFAE_B200_LIB"] = os.path.join(ROOT, "tools", "_build", "libnafae_b200_trace.so")
sys.path.insert(0, ROOT)
import numpy as np, torch
from nafae_b200 import synth, _C, parallel
from nafae_b200.pipeline import GroundingStep, capture_pipelined
import torch.distributed as dist

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
reserve = int(sys.argv[2]) if len(sys.argv) > 2 else 16
gate_head = len(sys.argv) > 3 and sys.argv[3] == "gate"
c = synth.CONFIGS[cfg]
rank, world, local = parallel.init_from_env()   # torchrun for world > 1 (adds the all-reduce branch)
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
steps = []
for i in range(2):
    st = GroundingStep(c["Na"], c["Ns"], c["Nb"], c["Ne"], c["D"], c["C"], c["H"], c["W"], c["n"],
                       pre_nms_topn=c["pre"], Delta=c["Delta"], vis_lam=c["vis_lam"], train=c["train"], device=dev,
                       tensor_cores=os.environ.get("TC") == "1")
    st.load(synth.make_batch(cfg, 1234 + i))
    st.run()
    steps.append(st)
torch.cuda.synchronize()
side = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
comm = torch.cuda.Stream(dev, priority=int(os.environ.get("COMM_PRIO", "-1")))
buckets = None
if world > 1:
    kw = {}
    if os.environ.get("AR_CTAS"):
        kw["num_ctas"] = int(os.environ["AR_CTAS"])
    mkw = dict(kw, cta_threads=int(os.environ["AR_THREADS"])) if os.environ.get("AR_THREADS") else kw
    buckets = [parallel.make_allreduce(parallel.trainable_grad_elems(), dev, kind=os.environ.get("AR_KIND", "auto"),
                                       peer_kw=kw, mc_kw=mkw) for _ in range(2)]
    for st, b in zip(steps, buckets):
        st.grad_word = b.views([(st.NQ, c["D"])])[0]
        b.launch()
    torch.cuda.synchronize()
_C.lib.nafae_set_reserved_sms(reserve)


def branch(j):
    if world <= 1:
        return None

    def br(cur):
        comm.wait_stream(cur)
        with torch.cuda.stream(comm):
            if not os.environ.get("NO_GATE"):
                steps[j].wait_gate(1)
            buckets[j].launch()
        return comm
    return br


gs = [capture_pipelined(steps[j], steps[1 - j], side, branch(j), gate_head=gate_head) for j in range(2)]
REC = np.dtype([("t0", "<u8"), ("t1", "<u8"), ("kernel", "<i4"), ("cta", "<i4"), ("smid", "<i4"),
                ("a", "<i4"), ("b", "<i4"), ("pad", "<i4")])
readers = []
for nm in ("nafae_debug_cta_trace_roi_align", "nafae_debug_cta_trace_ground", "nafae_debug_cta_trace_nms",
           "nafae_debug_cta_trace_allreduce", "nafae_debug_cta_trace_runtime"):
    fn = getattr(_C.lib, nm)
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    fn.restype = ctypes.c_int
    readers.append(fn)


def read(reset):
    out = []
    for fn in readers:
        buf = np.zeros(1 << 15, REC)
        n = fn(buf.ctypes.data, len(buf), int(reset))
        out.append(buf[:n])
    return np.concatenate(out)


for i in range(40):
    gs[i & 1].replay()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
read(True)
N = 6
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(N):
    gs[i & 1].replay()
e1.record()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
if rank != 0:
    if buckets:
        for b in buckets:
            b.close()
    sys.exit(0)
print("traced build: %.1f us/step over %d replays (reserve %d, gate_head %s)" % (e0.elapsed_time(e1) / N * 1e3, N, reserve, gate_head))
r = read(True)
names = {1: "roi_align", 2: "prop_tail", 3: "ground_fwd", 4: "ground_bwd", 5: "allreduce", 6: "gate_wait"}
# replays are separated by idle gaps (a graph replay starts after the previous one has fully ended)
r = np.sort(r, order="t0")
segs, start, tmax = [], 0, 0
for i in range(len(r)):
    if i > start and int(r["t0"][i]) > tmax:
        segs.append((start, i))
        start = i
    tmax = max(tmax, int(r["t1"][i]))
segs.append((start, len(r)))
print("segments (replays):", len(segs), [b - a for a, b in segs])
gaps = [(int(r["t0"][segs[i + 1][0]]) - int(r["t1"][segs[i][0]:segs[i][1]].max())) / 1e3 for i in range(len(segs) - 1)]
print("idle gap between consecutive replays (last CTA end -> first CTA start), us:", ["%.1f" % g for g in gaps])
for (sa, sb) in segs[-2:]:
    seg = r[sa:sb]
    base = int(seg["t0"].min())
    print("--- replay: %d CTAs, span %.1f us (times in us from the first CTA start)" % (len(seg), (int(seg["t1"].max()) - base) / 1e3))
    for k in (2, 1, 3, 4, 6, 5):
        x = seg[seg["kernel"] == k]
        if len(x) == 0:
            continue
        t0 = (x["t0"].astype(np.int64) - base) / 1e3
        t1 = (x["t1"].astype(np.int64) - base) / 1e3
        print("%-11s ctas %4d  SMs %3d  start min %6.1f p50 %6.1f p90 %6.1f max %6.1f | end min %6.1f p50 %6.1f p90 %6.1f max %6.1f | life p50 %5.1f max %5.1f"
              % (names[k], len(x), len(np.unique(x["smid"])), t0.min(), np.percentile(t0, 50), np.percentile(t0, 90), t0.max(),
                 t1.min(), np.percentile(t1, 50), np.percentile(t1, 90), t1.max(), np.percentile(t1 - t0, 50), (t1 - t0).max()))
        if k == 1:
            print("            items/CTA min %d p50 %d max %d (total %d), RoI passes/CTA min %d max %d; starters later than 5 us: %d"
                  % (x["a"].min(), np.percentile(x["a"], 50), x["a"].max(), x["a"].sum(), x["b"].min(), x["b"].max(), (t0 > 5).sum()))
            late = x[t0 > 5]
            if len(late):
                lt0 = (late["t0"].astype(np.int64) - base) / 1e3
                print("            late: start p50 %.1f max %.1f, their items p50 %d" % (np.percentile(lt0, 50), lt0.max(), np.percentile(late["a"], 50)))
        if k == 5:
            print("            phases per CTA (us since its start, p50 / max): first barrier passed %.1f / %.1f, data done %.1f / %.1f, exit %.1f / %.1f; co-resident with RoIAlign CTAs on %d SMs"
                  % (np.percentile(x["a"], 50) / 1e3, x["a"].max() / 1e3, np.percentile(x["b"], 50) / 1e3, x["b"].max() / 1e3,
                     np.percentile(t1 - t0, 50), (t1 - t0).max(),
                     len(np.intersect1d(x["smid"], seg[seg["kernel"] == 1]["smid"]))))
        if k in (3, 4):  # long-lived CTAs of the head kernels: where and when
            long_ = x[(t1 - t0) > 4]
            lt0 = (long_["t0"].astype(np.int64) - base) / 1e3
            lt1 = (long_["t1"].astype(np.int64) - base) / 1e3
            print("            CTAs living > 4 us: %d on %d SMs, start p50 %.1f, end p50 %.1f max %.1f" % (
                len(long_), len(np.unique(long_["smid"])), np.percentile(lt0, 50) if len(long_) else 0,
                np.percentile(lt1, 50) if len(long_) else 0, lt1.max() if len(long_) else 0))
if buckets:
    for b in buckets:
        b.close()
