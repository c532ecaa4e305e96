#!/bin/bash
# usage (on an 8-GPU box): tools/gpu_mgpu8b.sh OUTDIR -- DP step with the tcgen05 head at N=8 / 4: all-reduce placement
OUT=${1:-gpurun_out/mgpu8b}
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
run 8 "tc | peer V1 x16 gated, reserve 32" --tensor-cores 1 --allreduce peer
run 8 "tc | peer V1 x16 ungated, reserve 32" --tensor-cores 1 --allreduce peer --no-gate
run 8 "tc | peer V1 x16 ungated, reserve 24" --tensor-cores 1 --allreduce peer --no-gate --reserve-sms 24
run 8 "tc | mc 8x512 ungated, reserve 24" --tensor-cores 1 --allreduce multicast --ar-ctas 8 --no-gate --reserve-sms 24
run 8 "tc | mc 8x512 ungated, reserve 16" --tensor-cores 1 --allreduce multicast --ar-ctas 8 --no-gate --reserve-sms 16
run 8 "tc | mc 16x512 gated, reserve 24" --tensor-cores 1 --allreduce multicast --ar-ctas 16 --reserve-sms 24
run 4 "tc | peer V1 x16 ungated, reserve 32" --tensor-cores 1 --allreduce peer --no-gate
run 4 "tc | mc 8x512 ungated, reserve 24" --tensor-cores 1 --allreduce multicast --ar-ctas 8 --no-gate --reserve-sms 24
echo "== timeline N=8: tc head, peer x16 ungated" | tee -a $OUT/summary.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
TC=1 AR_KIND=peer NO_GATE=1 timeout 200 $TR tools/timeline.py cfg2 32 > $OUT/timeline_peer8.txt 2>&1
grep -vE "^\*|OMP|^$|NCCL version|W1017" $OUT/timeline_peer8.txt | tail -30 | cut -c1-250 | tee -a $OUT/summary.txt
