#!/bin/bash
OUT=${1:-gpurun_out/iter12}
mkdir -p $OUT
echo "== pytest -m gpu (all, no -x)" | tee $OUT/summary.txt
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/summary.txt; tail -6 $OUT/pytest.log | tee -a $OUT/summary.txt
run() {
  label=$1; shift
  timeout 200 python bench.py --steps 1600 --warmup 48 --no-e2e --no-cpu-baseline "$@" > $OUT/b.json 2>$OUT/b.err
  python - "$label" $OUT <<'PY' | tee -a $OUT/summary.txt
import json, sys
out = sys.argv[2]
try:
    d = json.loads(open(out + '/b.json').read().strip().splitlines()[-1])
    r = d["roofline"]
    print("%-40s %8.0f seg/s %6.1f us/step  align %5.1f us (%d CTAs, %d reserved) frac %.3f step_frac %.3f launches/step %.1f" % (
        sys.argv[1], d["value"], d["ms_per_step"] * 1e3, r["kernel_us"], r["kernel_grid_sms"], r["reserved_sms"], r["frac"], r["step_frac"], d["gpu_launches"] / d["steps"]))
    open(out + '/lines.jsonl', 'a').write(json.dumps({"label": sys.argv[1], "line": d}) + "\n")
except Exception as e:
    print(sys.argv[1], "failed", e, open(out + '/b.err').read()[-1200:])
PY
}
run "exact, 8 steps/graph (default)"
run "exact, 4 steps/graph" --steps-per-graph 4
run "exact, 16 steps/graph" --steps-per-graph 16
run "exact, 8 steps/graph, reserve 12" --reserve-sms 12
run "exact, 8 steps/graph, reserve 24" --reserve-sms 24
run "exact, 8 steps/graph, fma head" --tensor-cores 0
run "joined, 1 step/graph (old default)" --schedule joined
run "cfg4 exact" --cfg cfg4
run "cfg2_real exact" --cfg cfg2_real
