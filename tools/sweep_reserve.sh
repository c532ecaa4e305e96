#!/bin/bash
# Dev tool: sweep the SM reservation of the pipelined schedule.
for r in "$@"; do
  timeout 200 python bench.py --steps 300 --warmup 20 --no-e2e --no-cpu-baseline --reserve-sms $r > /tmp/b_$r.json 2>/tmp/b_$r.err
  python - "$r" <<'PY'
import json, sys
r = sys.argv[1]
try:
    d = json.loads(open('/tmp/b_%s.json' % r).read().strip().splitlines()[-1])
    print("reserve %3s: %8.0f seg/s  %6.1f us/step  align %5.1f us  frac %.3f  step_frac %.3f" % (
        r, d["value"], d["ms_per_step"] * 1e3, d["roofline"]["kernel_us"], d["roofline"]["frac"], d["roofline"]["step_frac"]))
except Exception as e:
    print("reserve", r, "failed", e, open('/tmp/b_%s.err' % r).read()[-500:])
PY
done
