#!/bin/bash
OUT=${1:-gpurun_out/exact8}
mkdir -p $OUT
: > $OUT/summary.txt
run() {
  n=$1; label=$2; shift; shift
  TRn="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541"
  timeout 150 $TRn bench.py --gpus $n --steps 2000 --warmup 48 --no-e2e --no-cpu-baseline "$@" > $OUT/b.json 2>$OUT/b.err
  python - "$label" $OUT <<'PY' | tee -a $OUT/summary.txt
import json, sys
out = sys.argv[2]
try:
    d = json.loads(open(out + '/b.json').read().strip().splitlines()[-1])
    print("%-30s N=%d %8.0f seg/s %6.1f us/step  align %5.1f us (%d CTAs, frac %.3f) kind=%s identical=%s launches/step %.1f" % (
        sys.argv[1], d["n_gpus"], d["value"], d["ms_per_step"] * 1e3, d["roofline"]["kernel_us"],
        d["roofline"]["kernel_grid_sms"], d["roofline"]["frac"], d["config"].get("allreduce_kind"), d.get("replicas_identical"),
        d["gpu_launches"] / d["steps"]))
    open(out + '/lines.jsonl', 'a').write(json.dumps({"label": sys.argv[1], "line": d}) + "\n")
except Exception as e:
    print(sys.argv[1], "failed", e, open(out + '/b.err').read()[-1200:])
PY
}
run 8 "exact, default (NVLS)"
run 4 "exact, NVLS" --allreduce multicast
run 2 "exact, NVLS" --allreduce multicast
