#!/bin/bash
# usage: sweep_comm.sh NGPU comm_sms...
N=$1; shift
for c in "$@"; do
  NAFAE_COMM_SMS=${c%%:*} NAFAE_AR_CTAS=${c##*:} timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 300 --warmup 20 --no-e2e $EXTRA > /tmp/c_$c.json 2>/tmp/c_$c.err
  python - "$c" <<'PY'
import json, sys
c = sys.argv[1]
try:
    d = json.loads(open('/tmp/c_%s.json' % c).read().strip().splitlines()[-1])
    print("comm_sms %3s: N=%d %8.0f seg/s  %6.1f us/step  align %5.1f us on %d SMs" % (
        c, d["n_gpus"], d["value"], d["ms_per_step"] * 1e3, d["roofline"]["kernel_us"], d["roofline"]["kernel_grid_sms"]))
except Exception as e:
    print("comm_sms", c, "failed", e, open('/tmp/c_%s.err' % c).read()[-800:])
PY
done
