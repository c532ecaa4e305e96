#!/bin/bash
# usage (on an N-GPU box): tools/gpu_mgpu_prio.sh NGPU OUTDIR -- all-reduce placement / priority sweep
N=${1:-2}
OUT=${2:-gpurun_out/mgpu_prio}
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
: > $OUT/summary.txt
run() {  # label, env..., -- extra bench args
  label=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 200 $TR bench.py --gpus $N --steps 1500 --warmup 50 --no-e2e --no-cpu-baseline "$@" > $OUT/b.json 2>$OUT/b.err
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
run "mc 16x256 co-resident, ungated, high prio" -- --allreduce multicast --ar-ctas 16 --ar-threads 256 --no-gate
run "mc 32x256 co-resident, ungated, high prio" -- --allreduce multicast --ar-ctas 32 --ar-threads 256 --no-gate
run "mc 16x256 co-resident, ungated, prio 0" -- --allreduce multicast --ar-ctas 16 --ar-threads 256 --no-gate --comm-priority 0
run "mc 16x256 co-resident, gated, high prio" -- --allreduce multicast --ar-ctas 16 --ar-threads 256
run "mc 8x512 / 4 SMs, gated, high prio" -- --allreduce multicast --ar-ctas 8 --comm-sms 4
run "mc 8x512 / 8 SMs, ungated, high prio" -- --allreduce multicast --ar-ctas 8 --comm-sms 8 --no-gate
run "peer V1 x16 / 16 SMs, gated, high prio" -- --allreduce peer
run "peer V1 x16 / 16 SMs, gated, prio 0 (old default)" -- --allreduce peer --comm-priority 0
echo "== timeline: mc 16x256 co-resident ungated high prio" | tee -a $OUT/summary.txt
AR_KIND=multicast AR_CTAS=16 AR_THREADS=256 NO_GATE=1 timeout 200 $TR tools/timeline.py cfg2 16 > $OUT/timeline_a.txt 2>&1
grep -vE "^\*|OMP|^$|NCCL version|W1017" $OUT/timeline_a.txt | tail -30 | cut -c1-250 | tee -a $OUT/summary.txt
