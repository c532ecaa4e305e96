#!/bin/bash
# usage (on the GPU box): tools/gpu_iter9.sh OUTDIR -- round-2 single-GPU validation with the final defaults
OUT=${1:-gpurun_out/iter9}
mkdir -p $OUT
echo "== pytest -m gpu (all, no -x)" | tee $OUT/summary.txt
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/summary.txt; tail -8 $OUT/pytest.log | tee -a $OUT/summary.txt
echo "== smoke" | tee -a $OUT/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee -a $OUT/summary.txt
echo "== operator lines" | tee -a $OUT/summary.txt
timeout 300 python tools/time_ops.py 2>&1 | tail -12 | tee -a $OUT/summary.txt
timeout 300 python tools/time_ops.py --cfg cfg2_real 2>&1 | tail -12 | tee -a $OUT/summary.txt
echo "== default bench line (driver form)" | tee -a $OUT/summary.txt
timeout 600 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err
tail -c 3000 $OUT/bench_default.json | tee -a $OUT/summary.txt
run() {
  label=$1; shift
  timeout 200 python bench.py --steps 1500 --warmup 50 --no-e2e --no-cpu-baseline "$@" > $OUT/b.json 2>$OUT/b.err
  python - "$label" $OUT <<'PY' | tee -a $OUT/summary.txt
import json, sys
out = sys.argv[2]
try:
    d = json.loads(open(out + '/b.json').read().strip().splitlines()[-1])
    if "roofline" in d and d["roofline"]:
        r = d["roofline"]
        print("%-40s %8.0f seg/s %6.1f us/step  align %5.1f us (%d CTAs, %d reserved) frac %.3f step_frac %.3f" % (
            sys.argv[1], d["value"], d["ms_per_step"] * 1e3, r["kernel_us"], r["kernel_grid_sms"], r["reserved_sms"], r["frac"], r["step_frac"]))
    else:
        print("%-40s %8.0f seg/s %6.1f us/step parity %s" % (sys.argv[1], d["value"], d["ms_per_step"] * 1e3, d.get("parity")))
    open(out + '/lines.jsonl', 'a').write(json.dumps({"label": sys.argv[1], "line": d}) + "\n")
except Exception as e:
    print(sys.argv[1], "failed", e, open(out + '/b.err').read()[-800:])
PY
}
echo "== other configs" | tee -a $OUT/summary.txt
run "cfg2 default (tcgen05 head)"
run "cfg2 fma head" --tensor-cores 0
run "cfg2 sequential" --no-pipeline
run "cfg4" --cfg cfg4
run "cfg4 fma head" --cfg cfg4 --tensor-cores 0
run "cfg2_real" --cfg cfg2_real
run "cfg5" --cfg cfg5
echo "== ncu" | tee -a $OUT/summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches.csv python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:align_pool_fwd_slab -s 12 -c 2 -o $OUT/prof_slab python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --no-pipeline > $OUT/ncu_slab.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:align_avg_bwd -c 3 -o $OUT/prof_bwd python tools/time_ops.py > $OUT/ncu_bwd.log 2>&1
ls -la $OUT | tee -a $OUT/summary.txt
