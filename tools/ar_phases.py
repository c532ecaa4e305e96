"""Dev tool (torchrun, trace build): where the time of one standalone all-reduce goes, per CTA:
launch -> first cross-GPU barrier passed -> data phase done -> (fence / drain + last barrier) -> exit.
Graph replays back to back, like tests/_mgpu_worker.py times them."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["NAFAE_B200_LIB"] = os.path.join(ROOT, "tools", "_build", "libnafae_b200_trace.so")
sys.path.insert(0, ROOT)
import numpy as np, torch
import torch.distributed as dist
from nafae_b200 import _C, parallel

rank, world, local = parallel.init_from_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
N = parallel.trainable_grad_elems()
REC = np.dtype([("t0", "<u8"), ("t1", "<u8"), ("kernel", "<i4"), ("cta", "<i4"), ("smid", "<i4"),
                ("a", "<i4"), ("b", "<i4"), ("pad", "<i4")])
fn = _C.lib.nafae_debug_cta_trace_allreduce
fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
fn.restype = ctypes.c_int


def read():
    buf = np.zeros(1 << 15, REC)
    n = fn(buf.ctypes.data, len(buf), 1)
    return buf[:n]


def run(name, ar, reps=20):
    ar.buf.normal_()
    torch.cuda.synchronize()
    dist.barrier()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        ar.launch()
    for _ in range(10):
        g.replay()
    torch.cuda.synchronize()
    dist.barrier()
    read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    r = read()
    us = e0.elapsed_time(e1) / reps * 1e3
    r = np.sort(r, order="t0")
    nc = ar.num_ctas
    per = [r[i * nc:(i + 1) * nc] for i in range(len(r) // nc)]
    life = np.array([(int(x["t1"].max()) - int(x["t0"].min())) / 1e3 for x in per])
    gap = np.array([(int(per[i + 1]["t0"].min()) - int(per[i]["t1"].max())) / 1e3 for i in range(len(per) - 1)])
    a = np.array([np.median(x["a"]) / 1e3 for x in per])
    b = np.array([np.median(x["b"]) / 1e3 for x in per])
    tot = np.array([np.median((x["t1"] - x["t0"]).astype(np.int64)) / 1e3 for x in per])
    line = ("%-28s rank %d: %.1f us/replay | kernel span %.1f | idle between replays %.1f | per CTA (median): "
            "barrier0 %.1f  data %.1f  drain+barrier2 %.1f" % (name, rank, us, np.median(life), np.median(gap),
                                                               np.median(a), np.median(b - a), np.median(tot - b)))
    lines = [None] * world
    dist.all_gather_object(lines, line)
    if rank == 0:
        print(lines[0], flush=True)
        print(lines[-1], flush=True)
    ar.close()


for ctas, thr in ((8, 512), (16, 512), (32, 512), (64, 256)):
    run("multicast %dx%d" % (ctas, thr), parallel.MulticastAllReduce(N, dev, num_ctas=ctas, cta_threads=thr))
for ctas in (8, 16, 32):
    run("peer V1 x%d" % ctas, parallel.PeerAllReduce(N, dev, num_ctas=ctas, cta_threads=0, variant=1))
# a 16x smaller bucket: what is left is the fixed cost
run("multicast 8x512, N/16", parallel.MulticastAllReduce(N // 16, dev, num_ctas=8, cta_threads=512))
run("peer V1 x8, N/16", parallel.PeerAllReduce(N // 16, dev, num_ctas=8, cta_threads=0, variant=1))
dist.destroy_process_group()
