#!/bin/bash
# usage (on the GPU box): tools/gpu_mgpu8.sh NGPU OUTDIR -- N-GPU validation + all-reduce sweeps at scale
N=${1:-8}
OUT=${2:-gpurun_out/mgpu8}
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
echo "== worker (correctness + standalone timing) N=$N" | tee $OUT/summary.txt
NAFAE_MGPU_TIME=1 timeout 600 $TR tests/_mgpu_worker.py > $OUT/worker.log 2>&1
echo "worker rc=$?" | tee -a $OUT/summary.txt
grep -E "allreduce|multicast|pipelined|HeadTrainer|FAIL|MGPU_OK|rror" $OUT/worker.log | tee -a $OUT/summary.txt
run() {  # label, env..., -- extra bench args
  label=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 240 $TR bench.py --gpus $N --steps 1500 --warmup 50 --no-e2e --no-cpu-baseline "$@" > $OUT/b.json 2>$OUT/b.err
  python - "$label" $OUT <<'PY' | tee -a $OUT/summary.txt
import json, sys
out = sys.argv[2]
try:
    d = json.loads(open(out + '/b.json').read().strip().splitlines()[-1])
    if "roofline" in d:
        print("%-46s N=%d %8.0f seg/s %6.1f us/step  align %5.1f us (%d CTAs) kind=%s identical=%s" % (
            sys.argv[1], d["n_gpus"], d["value"], d["ms_per_step"] * 1e3, d["roofline"]["kernel_us"],
            d["roofline"]["kernel_grid_sms"], d["config"].get("allreduce_kind"), d.get("replicas_identical")))
    else:
        print("%-46s N=%d %8.0f seg/s %6.1f us/step parity %s" % (sys.argv[1], d["n_gpus"], d["value"], d["ms_per_step"] * 1e3, d.get("parity")))
except Exception as e:
    print(sys.argv[1], "failed", e, open(out + '/b.err').read()[-800:])
PY
}
echo "== bench sweeps (N=$N)" | tee -a $OUT/summary.txt
run "multicast 16 ctas / 8 SMs (default)" -- --allreduce multicast
run "multicast 8 ctas / 4 SMs" -- --allreduce multicast --ar-ctas 8 --comm-sms 4
run "multicast 4 ctas / 2 SMs" -- --allreduce multicast --ar-ctas 4 --comm-sms 2
run "multicast 32 ctas / 16 SMs" -- --allreduce multicast --ar-ctas 32 --comm-sms 16
run "multicast 8 ctas / 0 comm SMs" -- --allreduce multicast --ar-ctas 8 --comm-sms 0
run "peer V1 x16 on 16 SMs" -- --allreduce peer
run "peer V0 x16 on 16 SMs" NAFAE_AR_VARIANT=0 -- --allreduce peer
run "cfg5 10000 segments" -- --cfg cfg5
