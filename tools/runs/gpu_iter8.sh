#!/bin/bash
OUT=${1:-gpurun_out/iter8}
mkdir -p $OUT
echo "== pytest -m gpu (all, no -x)" | tee $OUT/summary.txt
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/summary.txt; tail -12 $OUT/pytest.log | tee -a $OUT/summary.txt
echo "== operator lines" | tee -a $OUT/summary.txt
timeout 300 python tools/time_ops.py 2>&1 | tail -12 | tee -a $OUT/summary.txt
timeout 300 python tools/time_ops.py --cfg cfg2_real 2>&1 | tail -12 | tee -a $OUT/summary.txt
echo "== bridge" | tee -a $OUT/summary.txt
timeout 300 python tools/time_bridge.py 2>&1 | tail -10 | tee -a $OUT/summary.txt
run() {
  label=$1; shift
  timeout 200 python bench.py --steps 1500 --warmup 50 --no-e2e --no-cpu-baseline "$@" > $OUT/b.json 2>$OUT/b.err
  python - "$label" $OUT <<'PY' | tee -a $OUT/summary.txt
import json, sys
out = sys.argv[2]
try:
    d = json.loads(open(out + '/b.json').read().strip().splitlines()[-1])
    print("%-40s %8.0f seg/s %6.1f us/step  align %5.1f us (%d CTAs, %d reserved) frac %.3f step_frac %.3f" % (
        sys.argv[1], d["value"], d["ms_per_step"] * 1e3, d["roofline"]["kernel_us"],
        d["roofline"]["kernel_grid_sms"], d["roofline"]["reserved_sms"], d["roofline"]["frac"], d["roofline"]["step_frac"]))
except Exception as e:
    print(sys.argv[1], "failed", e, open(out + '/b.err').read()[-800:])
PY
}
echo "== N=1 step: FMA head vs tcgen05 head, reserved SMs" | tee -a $OUT/summary.txt
run "fma head, reserve 16 (default)"
run "tcgen05 head, reserve 16" --tensor-cores 1
run "tcgen05 head, reserve 12" --tensor-cores 1 --reserve-sms 12
run "tcgen05 head, reserve 8" --tensor-cores 1 --reserve-sms 8
run "tcgen05 head, reserve 20" --tensor-cores 1 --reserve-sms 20
run "tcgen05 head, reserve 16, 4 steps/graph" --tensor-cores 1 --steps-per-graph 4
run "fma head, reserve 16, 4 steps/graph" --steps-per-graph 4
echo "== timeline, tcgen05 head, reserve 16" | tee -a $OUT/summary.txt
TC=1 timeout 200 python tools/timeline.py cfg2 16 > $OUT/timeline_tc.txt 2>&1
tail -22 $OUT/timeline_tc.txt | cut -c1-250 | tee -a $OUT/summary.txt
