#!/bin/bash
# usage: tools/gpu_slab_ab.sh OUTDIR -- A/B of RoIAlign slab-kernel scheduling variants (1 GPU)
OUT=${1:-gpurun_out/slab_ab}
mkdir -p $OUT
echo "== pytest (roi_align, pipeline)" | tee $OUT/summary.txt
timeout 600 python -m pytest tests/test_roi_align.py tests/test_pipeline.py -m gpu -x -q > $OUT/pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/summary.txt; tail -4 $OUT/pytest.log | tee -a $OUT/summary.txt
run() {  # label, env..., -- extra bench args
  label=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 200 python bench.py --steps 1500 --warmup 50 --no-e2e --no-cpu-baseline "$@" > $OUT/b.json 2>$OUT/b.err
  python - "$label" $OUT <<'PY' | tee -a $OUT/summary.txt
import json, sys
out = sys.argv[2]
try:
    d = json.loads(open(out + '/b.json').read().strip().splitlines()[-1])
    r = d["roofline"]
    print("%-44s %8.0f seg/s %6.1f us/step  align %5.1f us (%d CTAs) frac %.3f step_frac %.3f" % (
        sys.argv[1], d["value"], d["ms_per_step"] * 1e3, r["kernel_us"], r["kernel_grid_sms"], r["frac"], r["step_frac"]))
except Exception as e:
    print(sys.argv[1], "failed", e, open(out + '/b.err').read()[-800:])
PY
}
B=tools/_build
run "dynamic (default) r16" --
run "dynamic r0 sequential" -- --no-pipeline
run "static split r16" -- --static-slab
run "static split r0 sequential" -- --static-slab --no-pipeline
run "scav3 r16" NAFAE_B200_LIB=$B/libnafae_b200_scav3.so --
run "scav3 r0 sequential" NAFAE_B200_LIB=$B/libnafae_b200_scav3.so -- --no-pipeline
run "dynamic r8" -- --reserve-sms 8
run "dynamic r12" -- --reserve-sms 12
run "dynamic r20" -- --reserve-sms 20
for extra in "$@"; do :; done
timeout 120 python tools/timeline.py cfg2 16 > $OUT/timeline_r16.txt 2>&1; tail -11 $OUT/timeline_r16.txt | tee -a $OUT/summary.txt
