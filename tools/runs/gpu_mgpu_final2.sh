#!/bin/bash
# usage: tools/gpu_mgpu_final2.sh NGPU OUTDIR -- worker + default bench lines at N
N=${1:-2}
OUT=${2:-gpurun_out/final2}
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
echo "== worker N=$N" | tee $OUT/summary.txt
timeout 500 $TR tests/_mgpu_worker.py > $OUT/worker.log 2>&1
echo "worker rc=$?" | tee -a $OUT/summary.txt
grep -E "pipelined|HeadTrainer|FAIL|MGPU_OK|rror" $OUT/worker.log | tee -a $OUT/summary.txt
run() {
  label=$1; shift
  timeout 200 $TR bench.py --gpus $N --steps 1500 --warmup 50 --no-e2e --no-cpu-baseline "$@" > $OUT/b.json 2>$OUT/b.err
  python - "$label" $OUT <<'PY' | tee -a $OUT/summary.txt
import json, sys
out = sys.argv[2]
try:
    d = json.loads(open(out + '/b.json').read().strip().splitlines()[-1])
    print("%-40s N=%d %8.0f seg/s %6.1f us/step  align %5.1f us (%d CTAs) kind=%s identical=%s launches/step %.1f" % (
        sys.argv[1], d["n_gpus"], d["value"], d["ms_per_step"] * 1e3, d["roofline"]["kernel_us"],
        d["roofline"]["kernel_grid_sms"], d["config"].get("allreduce_kind"), d.get("replicas_identical"), d["gpu_launches"] / d["steps"]))
    open(out + '/lines.jsonl', 'a').write(json.dumps({"label": sys.argv[1], "line": d}) + "\n")
except Exception as e:
    print(sys.argv[1], "failed", e, open(out + '/b.err').read()[-800:])
PY
}
run "default"
run "default, 1 step/graph" --steps-per-graph 1
run "default, 8 steps/graph" --steps-per-graph 8
