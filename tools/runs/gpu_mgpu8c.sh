#!/bin/bash
# usage (on an 8-GPU box): tools/gpu_mgpu8c.sh OUTDIR -- N=8: few-CTA NVLS all-reduce, multi-step graphs
OUT=${1:-gpurun_out/mgpu8c}
mkdir -p $OUT
: > $OUT/summary.txt
run() {  # N, label, extra bench args
  N=$1; label=$2; shift; shift
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
  timeout 200 $TR bench.py --gpus $N --steps 1500 --warmup 50 --no-e2e --no-cpu-baseline "$@" > $OUT/b.json 2>$OUT/b.err
  python - "$label" $OUT <<'PY' | tee -a $OUT/summary.txt
import json, sys
out = sys.argv[2]
try:
    d = json.loads(open(out + '/b.json').read().strip().splitlines()[-1])
    print("%-52s N=%d %8.0f seg/s %6.1f us/step  align %5.1f us (%d CTAs) kind=%s identical=%s" % (
        sys.argv[1], d["n_gpus"], d["value"], d["ms_per_step"] * 1e3, d["roofline"]["kernel_us"],
        d["roofline"]["kernel_grid_sms"], d["config"].get("allreduce_kind"), d.get("replicas_identical")))
    open(out + '/lines.jsonl', 'a').write(json.dumps({"label": sys.argv[1], "line": d}) + "\n")
except Exception as e:
    print(sys.argv[1], "failed", e, open(out + '/b.err').read()[-800:])
PY
}
run 8 "tc | mc 4x512 ungated, reserve 18" --allreduce multicast --ar-ctas 4 --no-gate --reserve-sms 18
run 8 "tc | mc 4x512 ungated, reserve 20" --allreduce multicast --ar-ctas 4 --no-gate --reserve-sms 20
run 8 "tc | mc 4x512 gated, reserve 20" --allreduce multicast --ar-ctas 4 --reserve-sms 20
run 8 "tc | mc 4x1024 ungated, reserve 20" --allreduce multicast --ar-ctas 4 --ar-threads 1024 --no-gate --reserve-sms 20
run 8 "tc | peer x16 gated, reserve 32, 4 steps/graph" --allreduce peer --steps-per-graph 4
run 8 "tc | mc 8x512 ungated, reserve 24, 4 steps/graph" --allreduce multicast --ar-ctas 8 --no-gate --reserve-sms 24 --steps-per-graph 4
