#!/bin/bash
# usage (on the GPU box): tools/gpu_single.sh OUTDIR [quick] -- 1-GPU validation: tests, bench sweeps, timeline, ncu
OUT=${1:-gpurun_out/single}
mkdir -p $OUT
echo "== pytest -m gpu" | tee $OUT/summary.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -15 $OUT/pytest.log | tee -a $OUT/summary.txt
run() {  # label -- extra bench args
  label=$1; shift; shift
  timeout 200 python bench.py --steps 2000 --warmup 50 --no-e2e --no-cpu-baseline "$@" > $OUT/b.json 2>$OUT/b.err
  python - "$label" $OUT <<'PY' | tee -a $OUT/summary.txt
import json, sys
out = sys.argv[2]
try:
    d = json.loads(open(out + '/b.json').read().strip().splitlines()[-1])
    r = d["roofline"]
    print("%-40s %8.0f seg/s %6.1f us/step  align %5.1f us (%d CTAs) frac %.3f step_frac %.3f" % (
        sys.argv[1], d["value"], d["ms_per_step"] * 1e3, r["kernel_us"], r["kernel_grid_sms"], r["frac"], r["step_frac"]))
except Exception as e:
    print(sys.argv[1], "failed", e, open(out + '/b.err').read()[-800:])
PY
}
echo "== bench sweeps (N=1)" | tee -a $OUT/summary.txt
run "pipelined reserve 16 (default)" --
run "pipelined reserve 0" -- --reserve-sms 0
run "pipelined reserve 4" -- --reserve-sms 4
run "pipelined reserve 8" -- --reserve-sms 8
run "pipelined reserve 12" -- --reserve-sms 12
run "pipelined reserve 24" -- --reserve-sms 24
run "sequential" -- --no-pipeline
run "cfg4 pipelined" -- --cfg cfg4
run "cfg2_real pipelined" -- --cfg cfg2_real
echo "== timeline" | tee -a $OUT/summary.txt
timeout 120 python tools/timeline.py cfg2 16 > $OUT/timeline_r16.txt 2>&1; tail -22 $OUT/timeline_r16.txt | tee -a $OUT/summary.txt
timeout 120 python tools/timeline.py cfg2 8 > $OUT/timeline_r8.txt 2>&1; tail -22 $OUT/timeline_r8.txt | tee -a $OUT/summary.txt
echo "== ncu" | tee -a $OUT/summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches.csv python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:align_pool_fwd_slab -s 12 -c 2 -o $OUT/prof_slab python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --no-pipeline > $OUT/ncu_slab.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ground_ -s 8 -c 4 -o $OUT/prof_ground python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --no-pipeline > $OUT/ncu_ground.log 2>&1
ls -la $OUT | tee -a $OUT/summary.txt
