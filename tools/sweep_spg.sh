#!/bin/bash
for s in "$@"; do
  timeout 200 python bench.py --steps 320 --warmup 20 --no-e2e --no-cpu-baseline --steps-per-graph $s > /tmp/s_$s.json 2>/tmp/s_$s.err
  python - "$s" <<'PY'
import json, sys
r = sys.argv[1]
try:
    d = json.loads(open('/tmp/s_%s.json' % r).read().strip().splitlines()[-1])
    print("steps/graph %3s: %8.0f seg/s  %6.1f us/step  step_frac %.3f" % (r, d["value"], d["ms_per_step"] * 1e3, d["roofline"]["step_frac"]))
except Exception as e:
    print("spg", r, "failed", e, open('/tmp/s_%s.err' % r).read()[-600:])
PY
done
