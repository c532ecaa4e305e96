"""Dev tool: which branch of the software-pipelined step graph is the critical path?
Captures the step graph with subsets of its branches (A = RoIAlign, B = proposal tail, C = head
fwd+bwd) and prints us/step for each subset, same SM reservation as bench.py."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nafae_b200 import synth, _C
from nafae_b200.pipeline import GroundingStep

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
reserve = int(sys.argv[2]) if len(sys.argv) > 2 else 16
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
c = synth.CONFIGS[cfg]
dev = torch.device("cuda:0")
steps = []
for i in range(2):
    st = GroundingStep(c["Na"], c["Ns"], c["Nb"], c["Ne"], c["D"], c["C"], c["H"], c["W"], c["n"],
                       pre_nms_topn=c["pre"], Delta=c["Delta"], vis_lam=c["vis_lam"], train=c["train"], device=dev)
    st.load(synth.make_batch(cfg, 1234 + i))
    st.run()
    steps.append(st)
torch.cuda.synchronize()
side = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
hi = torch.cuda.Stream(dev, priority=-1)
lo = [torch.cuda.Stream(dev, priority=0), torch.cuda.Stream(dev, priority=0)]
print("priority range", torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else "n/a")


def capture3(j, res, gated):
    from nafae_b200.pipeline import capture_pipelined
    _C.lib.nafae_set_reserved_sms(res)
    return capture_pipelined(steps[j], steps[1 - j], side, None, gate_head=gated)


def capture2(j, res, a_prio, a_first):
    """A on its own (optionally high-priority) stream, created first or last."""
    _C.lib.nafae_set_reserved_sms(res)
    a, n = steps[j], steps[1 - j]
    sa = hi if a_prio else lo[0]
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        cur = torch.cuda.current_stream()
        def doA():
            sa.wait_stream(cur)
            with torch.cuda.stream(sa):
                a.run_align()
        if a_first:
            doA()
        for st, fn in ((side[0], n.run_tail), (side[1], n.run_head)):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                fn()
        if not a_first:
            doA()
        for st in (sa, side[0], side[1]):
            cur.wait_stream(st)
    return g


def capture(sub, j, res):
    _C.lib.nafae_set_reserved_sms(res)
    a, n = steps[j], steps[1 - j]
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        cur = torch.cuda.current_stream()
        joined = []
        for key, st, fn in (("B", side[0], n.run_tail), ("C", side[1], n.run_head)):
            if key in sub:
                if sub == key:  # single branch: run on the capturing stream
                    fn()
                    continue
                st.wait_stream(cur)
                with torch.cuda.stream(st):
                    fn()
                joined.append(st)
        if "A" in sub:
            a.run_align()
        for st in joined:
            cur.wait_stream(st)
    return g


variants = [("A", 0), ("A", reserve), ("AC", reserve), ("ABC", reserve)]
variants += [("gate0", res) for res in (8, 24)] + [("gate1", res) for res in (16, 32)]
for sub, res in variants:
    if sub.startswith("gate"):
        gs = [capture3(j, res, sub[-1] == "1") for j in range(2)]
    elif sub.startswith("hi"):
        gs = [capture2(j, res, sub[2] == "1", sub[-1] == "1") for j in range(2)]
    else:
        gs = [capture(sub, j, res) for j in range(2)]
    for i in range(50):
        gs[i & 1].replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        gs[i & 1].replay()
    e1.record()
    torch.cuda.synchronize()
    print("branches %-10s reserve %2d : %7.2f us/step" % (sub, res, e0.elapsed_time(e1) / iters * 1e3))
