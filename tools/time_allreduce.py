"""Dev tool: time the 8.8 MB gradient-bucket all-reduce alone (run under torchrun)."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nafae_b200 import parallel
rank, world, local = parallel.init_from_env()
torch.cuda.set_device(local)
buf = torch.zeros(parallel.trainable_grad_elems(), device="cuda")
for _ in range(20):
    dist.all_reduce(buf, op=dist.ReduceOp.AVG)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    dist.all_reduce(buf, op=dist.ReduceOp.AVG)
for _ in range(10):
    g.replay()
torch.cuda.synchronize()
e0.record()
for _ in range(200):
    g.replay()
e1.record()
torch.cuda.synchronize()
if rank == 0:
    print("world %d NCCL_MAX_NCHANNELS=%s all-reduce 8.8MB: %.1f us" % (world, os.environ.get("NCCL_MAX_NCHANNELS"), e0.elapsed_time(e1) / 200 * 1e3))
dist.destroy_process_group()
