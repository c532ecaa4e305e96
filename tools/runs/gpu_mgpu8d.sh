#!/bin/bash
# usage (on an 8-GPU box): tools/gpu_mgpu8d.sh OUTDIR -- final default lines at N = 8, 4, 2 and the cfg5 sweep
OUT=${1:-gpurun_out/mgpu8d}
mkdir -p $OUT
: > $OUT/summary.txt
run() {  # N, label, extra bench args
  N=$1; label=$2; shift; shift
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
  timeout 200 $TR bench.py --gpus $N --steps 2000 --warmup 50 --no-e2e --no-cpu-baseline "$@" > $OUT/b.json 2>$OUT/b.err
  python - "$label" $OUT <<'PY' | tee -a $OUT/summary.txt
import json, sys
out = sys.argv[2]
try:
    d = json.loads(open(out + '/b.json').read().strip().splitlines()[-1])
    if d.get("roofline"):
        print("%-24s N=%d %8.0f seg/s %6.1f us/step  align %5.1f us (%d CTAs, frac %.3f) kind=%s identical=%s launches/step %.1f" % (
            sys.argv[1], d["n_gpus"], d["value"], d["ms_per_step"] * 1e3, d["roofline"]["kernel_us"],
            d["roofline"]["kernel_grid_sms"], d["roofline"]["frac"], d["config"].get("allreduce_kind"), d.get("replicas_identical"),
            d["gpu_launches"] / d["steps"]))
    else:
        print("%-24s N=%d %8.0f seg/s %6.1f us/step parity %s" % (sys.argv[1], d["n_gpus"], d["value"], d["ms_per_step"] * 1e3, d.get("parity")))
    open(out + '/lines.jsonl', 'a').write(json.dumps({"label": sys.argv[1], "line": d}) + "\n")
except Exception as e:
    print(sys.argv[1], "failed", e, open(out + '/b.err').read()[-1200:])
PY
}
run 8 "default"
run 4 "default"
run 2 "default"
run 8 "cfg5 sweep" --cfg cfg5
