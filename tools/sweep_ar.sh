#!/bin/bash
# usage: sweep_ar.sh NGPU  -- all-reduce variants standalone and inside the pipelined step (dev tool)
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
for cfg in "0 16" "0 32"; do
  set -- $cfg
  NAFAE_AR_THREADS=$1 NAFAE_AR_CTAS=$2 timeout 100 $TR tools/test_allreduce.py 2>&1 | grep -E "^world|rror" | sed "s/^/[threads $1 ctas $2] /"
  NAFAE_AR_VARIANT=1 NAFAE_AR_THREADS=$1 NAFAE_AR_CTAS=$2 timeout 100 $TR tools/test_allreduce.py 2>&1 | grep -E "^world|rror" | sed "s/^/[variant 1, threads $1 ctas $2] /"
done
run() {  # label, env..., -- extra bench args
  label=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 150 $TR bench.py --gpus $N --steps 1000 --warmup 30 --no-e2e --no-cpu-baseline "$@" > /tmp/ar.json 2>/tmp/ar.err
  python - "$label" <<'PY'
import json, sys
try:
    d = json.loads(open('/tmp/ar.json').read().strip().splitlines()[-1])
    print("%-44s N=%d %8.0f seg/s  %6.1f us/step  align %5.1f us on %d SMs" % (
        sys.argv[1], d["n_gpus"], d["value"], d["ms_per_step"] * 1e3, d["roofline"]["kernel_us"], d["roofline"]["kernel_grid_sms"]))
except Exception as e:
    print(sys.argv[1], "failed", e, open('/tmp/ar.err').read()[-600:])
PY
}
run "tma x16 on 16 SMs, gated" NAFAE_AR_THREADS=0 --
run "tma x16 variant 1 (3 slots, 2x chunks, parallel issue)" NAFAE_AR_THREADS=0 NAFAE_AR_VARIANT=1 --
run "tma x12 on 12 SMs, gated" NAFAE_AR_THREADS=0 NAFAE_COMM_SMS=12 --
run "tma x8 on 8 SMs, gated" NAFAE_AR_THREADS=0 NAFAE_COMM_SMS=8 --
