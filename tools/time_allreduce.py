"""Dev tool: time the 8.8 MB gradient-bucket all-reduce alone (run under torchrun, no graphs)."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 512 * 4096 + 512 + 512 * 200 + 512 + 1024
buf = torch.zeros(n, device="cuda")
for _ in range(20):
    dist.all_reduce(buf, op=dist.ReduceOp.AVG)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    dist.all_reduce(buf, op=dist.ReduceOp.AVG)
e1.record()
torch.cuda.synchronize()
if rank == 0:
    print("world %d NCCL_MAX_NCHANNELS=%s NCCL_ALGO=%s: all-reduce %.1f MB: %.1f us" % (
        world, os.environ.get("NCCL_MAX_NCHANNELS"), os.environ.get("NCCL_ALGO"), n * 4 / 1e6, e0.elapsed_time(e1) / 200 * 1e3), flush=True)
dist.destroy_process_group()
