#!/bin/bash
# usage (on an N-GPU box): tools/gpu_mgpu_tc.sh NGPU OUTDIR -- DP step with the tcgen05 head: all-reduce placement sweep
N=${1:-2}
OUT=${2:-gpurun_out/mgpu_tc}
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
: > $OUT/summary.txt
run() {  # label, extra bench args
  label=$1; shift
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
run "tc head | peer V1 x16 / 16 SMs gated" --tensor-cores 1 --allreduce peer
run "tc head | peer V1 x16 / 16 SMs ungated" --tensor-cores 1 --allreduce peer --no-gate
run "tc head | peer V1 x16 / 8 comm SMs gated" --tensor-cores 1 --allreduce peer --comm-sms 8
run "tc head | peer V1 x8 / 8 SMs gated" --tensor-cores 1 --allreduce peer --ar-ctas 8 --comm-sms 8
run "tc head | mc 8x512 / 4 SMs gated" --tensor-cores 1 --allreduce multicast --ar-ctas 8 --comm-sms 4
run "tc head | mc 8x512 / 8 SMs ungated" --tensor-cores 1 --allreduce multicast --ar-ctas 8 --comm-sms 8 --no-gate
run "tc head | mc 8x512 / 0 comm SMs gated" --tensor-cores 1 --allreduce multicast --ar-ctas 8 --comm-sms 0
run "tc head | mc 16x256 co-resident ungated" --tensor-cores 1 --allreduce multicast --ar-ctas 16 --ar-threads 256 --no-gate
run "tc head | mc 16x512 / 8 SMs gated" --tensor-cores 1 --allreduce multicast --ar-ctas 16 --comm-sms 8
echo "== timeline: tc head, peer x16 gated" | tee -a $OUT/summary.txt
TC=1 AR_KIND=peer timeout 200 $TR tools/timeline.py cfg2 32 > $OUT/timeline_peer.txt 2>&1
grep -vE "^\*|OMP|^$|NCCL version|W1017" $OUT/timeline_peer.txt | tail -16 | cut -c1-250 | tee -a $OUT/summary.txt
