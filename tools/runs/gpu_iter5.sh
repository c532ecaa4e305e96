#!/bin/bash
OUT=${1:-gpurun_out/iter5}
mkdir -p $OUT
echo "== pytest (gemm)" | tee $OUT/summary.txt
timeout 300 python -m pytest tests/test_gemm.py -m gpu -x -q > $OUT/pytest_gemm.log 2>&1
echo "pytest gemm rc=$?" | tee -a $OUT/summary.txt; tail -15 $OUT/pytest_gemm.log | tee -a $OUT/summary.txt
echo "== pytest (roi_align, pipeline)" | tee -a $OUT/summary.txt
timeout 600 python -m pytest tests/test_roi_align.py tests/test_pipeline.py -m gpu -x -q > $OUT/pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/summary.txt; tail -4 $OUT/pytest.log | tee -a $OUT/summary.txt
run() {  # label, env..., -- extra bench args
  label=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 200 python bench.py --no-e2e --no-cpu-baseline "$@" > $OUT/b.json 2>$OUT/b.err
  python - "$label" $OUT <<'PY' | tee -a $OUT/summary.txt
import json, sys
out = sys.argv[2]
try:
    d = json.loads(open(out + '/b.json').read().strip().splitlines()[-1])
    r = d.get("roofline")
    print("%-44s %8.0f seg/s %6.1f us/step  align %5.1f us (%d CTAs) frac %.3f step_frac %.3f" % (
        sys.argv[1], d["value"], d["ms_per_step"] * 1e3, r["kernel_us"], r["kernel_grid_sms"], r["frac"], r["step_frac"]))
except Exception as e:
    print(sys.argv[1], "failed", e, open(out + '/b.err').read()[-1500:])
PY
}
B=tools/_build
L="--steps 1500 --warmup 50"
run "tail3 (default) pipelined r16" -- $L
run "tail3 sequential" -- $L --no-pipeline
run "static-slab pipelined" -- $L --static-slab
run "static-slab sequential" -- $L --static-slab --no-pipeline
run "tail6 pipelined r16" NAFAE_B200_LIB=$B/libnafae_b200_tail6.so -- $L
run "tail6 sequential" NAFAE_B200_LIB=$B/libnafae_b200_tail6.so -- $L --no-pipeline
run "r1 slab pipelined" NAFAE_B200_LIB=$B/libnafae_b200_r1slab.so -- $L
run "r1 slab sequential" NAFAE_B200_LIB=$B/libnafae_b200_r1slab.so -- $L --no-pipeline
run "tail3 pipelined r12" -- $L --reserve-sms 12
run "tail3 pipelined r8" -- $L --reserve-sms 8
timeout 120 python tools/timeline.py cfg2 16 > $OUT/timeline_r16.txt 2>&1; tail -12 $OUT/timeline_r16.txt | tee -a $OUT/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:align_pool_fwd_slab -s 12 -c 2 -o $OUT/prof_slab python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --no-pipeline > $OUT/ncu_slab.log 2>&1
