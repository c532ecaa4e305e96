"""Dev/validation tool (torchrun): peer-memory all-reduce vs NCCL, eager + CUDA-graph, with timing."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nafae_b200 import parallel
rank, world, local = parallel.init_from_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
n = int(os.environ.get("N_FLOATS", parallel.trainable_grad_elems()))
n = (n + 4 * world - 1) // (4 * world) * (4 * world)
ar = parallel.PeerAllReduce(n, dev)
torch.manual_seed(100 + rank)
ok = True
for it in range(3):
    x = torch.randn(n, device=dev)
    ar.buf.copy_(x)
    ref = x.clone()
    dist.all_reduce(ref, op=dist.ReduceOp.SUM)
    ref /= world
    torch.cuda.synchronize(); dist.barrier()
    ar.launch()
    torch.cuda.synchronize()
    err = (ar.buf - ref).abs().max().item()
    # every replica must hold identical bits
    chk = ar.buf.double().sum().reshape(1).clone()
    lst = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(lst, chk)
    same = all(torch.equal(lst[0], t) for t in lst)
    ok = ok and err < 1e-5 and same
    if rank == 0:
        print("iter %d max|err| vs NCCL %.3e, replicas identical: %s" % (it, err, same), flush=True)
# graph capture + timing
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    ar.launch()
for _ in range(10):
    g.replay()
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    g.replay()
e1.record()
torch.cuda.synchronize()
if rank == 0:
    print("world %d peer all-reduce %.1f MB, %d CTAs: %.1f us (graph replay)  ok=%s" % (
        world, n * 4 / 1e6, ar.num_ctas, e0.elapsed_time(e1) / 200 * 1e3, ok), flush=True)
ar.close()
if rank == 0 and os.environ.get("P2P_BW") and torch.cuda.device_count() > 1:
    a = torch.empty(64 << 20, dtype=torch.uint8, device="cuda:0")
    b = torch.empty(64 << 20, dtype=torch.uint8, device="cuda:1")
    for _ in range(3):
        a.copy_(b)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        a.copy_(b)
    e1.record()
    torch.cuda.synchronize()
    print("p2p copy cuda:1 -> cuda:0, 64 MiB: %.0f GB/s" % (10 * (64 << 20) / (e0.elapsed_time(e1) * 1e-3) / 1e9), flush=True)
dist.destroy_process_group()
