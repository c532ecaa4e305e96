#!/bin/bash
OUT=${1:-gpurun_out/iter3}
mkdir -p $OUT
echo "== pytest" | tee $OUT/summary.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/summary.txt; tail -6 $OUT/pytest.log | tee -a $OUT/summary.txt
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
    if r:
        print("%-44s %8.0f seg/s %6.1f us/step  align %5.1f us (%d CTAs) frac %.3f step_frac %.3f clocks %s" % (
            sys.argv[1], d["value"], d["ms_per_step"] * 1e3, r["kernel_us"], r["kernel_grid_sms"], r["frac"], r["step_frac"], d["clocks"]["sm_mhz"]))
    else:
        print("%-44s %8.0f seg/s %6.1f us/step parity %s" % (sys.argv[1], d["value"], d["ms_per_step"] * 1e3, d.get("parity")))
except Exception as e:
    print(sys.argv[1], "failed", e, open(out + '/b.err').read()[-1500:])
PY
}
B=tools/_build
L="--steps 1500 --warmup 50"
S="--steps 20 --warmup 5"
run "v4 long" -- $L
run "v4 short(20)" -- $S
run "v4 sequential long" -- $L --no-pipeline
run "v4 sequential short" -- $S --no-pipeline
run "r1 slab long" NAFAE_B200_LIB=$B/libnafae_b200_r1slab.so -- $L
run "r1 slab short(20)" NAFAE_B200_LIB=$B/libnafae_b200_r1slab.so -- $S
run "r1 slab sequential long" NAFAE_B200_LIB=$B/libnafae_b200_r1slab.so -- $L --no-pipeline
run "r1 slab sequential short" NAFAE_B200_LIB=$B/libnafae_b200_r1slab.so -- $S --no-pipeline
run "bwd256 long" NAFAE_B200_LIB=$B/libnafae_b200_bwd256.so -- $L
run "bwd256 short" NAFAE_B200_LIB=$B/libnafae_b200_bwd256.so -- $S
run "bwd256 long r12" NAFAE_B200_LIB=$B/libnafae_b200_bwd256.so -- $L --reserve-sms 12
run "cfg5 2000 segments" -- --cfg cfg5 --segments 2000
timeout 120 python tools/timeline.py cfg2 16 > $OUT/timeline_r16.txt 2>&1; tail -12 $OUT/timeline_r16.txt | tee -a $OUT/summary.txt
NAFAE_B200_LIB= timeout 600 ncu --set full --clock-control none --import-source on -k regex:align_pool_fwd_slab -s 12 -c 2 -o $OUT/prof_slab python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --no-pipeline > $OUT/ncu_slab.log 2>&1
ls $OUT | tee -a $OUT/summary.txt
