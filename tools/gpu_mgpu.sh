#!/bin/bash
# usage (on the GPU box): tools/gpu_mgpu.sh NGPU OUTDIR -- multi-GPU validation + all-reduce sweeps
N=${1:-2}
OUT=${2:-gpurun_out/mgpu}
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo "== NCCL NVLS probe" | tee $OUT/summary.txt
NCCL_DEBUG=INFO timeout 120 $TR tools/time_allreduce.py 2>&1 | grep -iE "nvls|world" | head -8 | tee -a $OUT/summary.txt
echo "== worker (correctness + standalone timing)" | tee -a $OUT/summary.txt
NAFAE_MGPU_TIME=1 timeout 600 $TR tests/_mgpu_worker.py > $OUT/worker.log 2>&1
echo "worker rc=$?" | tee -a $OUT/summary.txt
grep -E "allreduce|multicast|pipelined|HeadTrainer|FAIL|MGPU_OK|rror" $OUT/worker.log | tee -a $OUT/summary.txt
run() {  # label, env..., -- extra bench args
  label=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 200 $TR bench.py --gpus $N --steps 1500 --warmup 50 --no-e2e --no-cpu-baseline "$@" > $OUT/b.json 2>$OUT/b.err
  python - "$label" $OUT <<'PY' | tee -a $OUT/summary.txt
import json, sys
out = sys.argv[2]
try:
    d = json.loads(open(out + '/b.json').read().strip().splitlines()[-1])
    print("%-46s N=%d %8.0f seg/s %6.1f us/step  align %5.1f us (%d CTAs) kind=%s identical=%s" % (
        sys.argv[1], d["n_gpus"], d["value"], d["ms_per_step"] * 1e3, d["roofline"]["kernel_us"],
        d["roofline"]["kernel_grid_sms"], d.get("run", {}).get("allreduce_kind"), d.get("replicas_identical")))
except Exception as e:
    print(sys.argv[1], "failed", e, open(out + '/b.err').read()[-800:])
PY
}
echo "== bench sweeps (N=$N)" | tee -a $OUT/summary.txt
run "auto (multicast if available) default" --
run "multicast 8 ctas / 4 SMs" -- --allreduce auto --ar-ctas 8 --comm-sms 4
run "multicast 16 ctas / 8 SMs, ungated" -- --no-gate
run "multicast 32 ctas / 16 SMs" -- --ar-ctas 32 --comm-sms 16
run "multicast 16 ctas, 0 comm SMs (shares head SMs)" -- --comm-sms 0
run "peer bulk-copy x16 on 16 SMs" -- --allreduce peer
run "nccl" -- --nccl-allreduce
