#!/bin/bash
# usage (on an 8-GPU box): tools/gpu_mgpu8.sh OUTDIR -- world-8 correctness, all-reduce sweeps and the
# 1/2/4/8 scaling table measured on one box
OUT=${1:-gpurun_out/mgpu8}
mkdir -p $OUT
tr() { echo "python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29541"; }
echo "== worker (correctness + standalone timing) N=8" | tee $OUT/summary.txt
NAFAE_MGPU_TIME=1 timeout 420 $(tr 8) tests/_mgpu_worker.py > $OUT/worker.log 2>&1
echo "worker rc=$?" | tee -a $OUT/summary.txt
grep -E "allreduce|multicast|pipelined|HeadTrainer|FAIL|MGPU_OK|rror" $OUT/worker.log | tee -a $OUT/summary.txt
run() {  # N, label, env..., -- extra bench args
  N=$1; label=$2; shift; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  if [ $N -gt 1 ]; then L=$(tr $N); else L=python; fi
  env "${envs[@]}" timeout 200 $L bench.py --gpus $N --steps 1500 --warmup 50 --no-e2e --no-cpu-baseline "$@" > $OUT/b.json 2>$OUT/b.err
  python - "$label" $OUT <<'PY' | tee -a $OUT/summary.txt
import json, sys
out = sys.argv[2]
try:
    d = json.loads(open(out + '/b.json').read().strip().splitlines()[-1])
    if "roofline" in d:
        print("%-40s N=%d %8.0f seg/s %6.1f us/step  align %5.1f us (%d CTAs) kind=%s identical=%s" % (
            sys.argv[1], d["n_gpus"], d["value"], d["ms_per_step"] * 1e3, d["roofline"]["kernel_us"],
            d["roofline"]["kernel_grid_sms"], d["config"].get("allreduce_kind"), d.get("replicas_identical")))
    else:
        print("%-40s N=%d %8.0f seg/s %6.1f us/step parity %s" % (sys.argv[1], d["n_gpus"], d["value"], d["ms_per_step"] * 1e3, d.get("parity")))
    open(out + '/lines.jsonl', 'a').write(json.dumps({"label": sys.argv[1], "line": d}) + "\n")
except Exception as e:
    print(sys.argv[1], "failed", e, open(out + '/b.err').read()[-800:])
PY
}
echo "== bench (one box)" | tee -a $OUT/summary.txt
run 8 "default (auto)" --
run 8 "multicast 8 ctas / 4 SMs" -- --allreduce multicast --ar-ctas 8 --comm-sms 4
run 8 "multicast 32 ctas / 16 SMs" -- --allreduce multicast --ar-ctas 32 --comm-sms 16
run 8 "peer V1 x16 on 16 SMs" -- --allreduce peer
run 4 "default (auto)" --
run 4 "peer V1 x16 on 16 SMs" -- --allreduce peer
run 2 "default (auto)" --
run 1 "default" --
run 8 "cfg5 sweep" -- --cfg cfg5
echo "== timeline N=8 (default all-reduce)" | tee -a $OUT/summary.txt
timeout 200 $(tr 8) tools/timeline.py cfg2 24 > $OUT/timeline8.txt 2>&1
tail -40 $OUT/timeline8.txt | tee -a $OUT/summary.txt
