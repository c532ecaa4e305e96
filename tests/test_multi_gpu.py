"""Multi-rank GPU tests: spawn one process per GPU with torch.distributed.run and let
tests/_mgpu_worker.py check every all-reduce variant, the pipelined data-parallel step and the
HeadTrainer replicas.  Skipped on boxes with a single GPU (the world_size-2 gloo tests in
test_parallel_cpu.py cover the host logic there)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_worker(nproc, extra_env=None, timeout=600):
    env = dict(os.environ)
    env.update(extra_env or {})
    env.setdefault("MASTER_ADDR", "127.0.0.1")
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "_mgpu_worker.py")]
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)


@pytest.mark.gpu
@pytest.mark.timeout(900)
def test_allreduce_variants_dp_step_and_trainer_on_all_gpus():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    out = _run_worker(world)
    sys.stdout.write(out.stdout[-4000:])
    assert out.returncode == 0 and "MGPU_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
