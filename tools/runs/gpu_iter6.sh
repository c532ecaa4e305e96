#!/bin/bash
OUT=${1:-gpurun_out/iter6}
mkdir -p $OUT
echo "== pytest -m gpu (all)" | tee $OUT/summary.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/summary.txt; tail -12 $OUT/pytest.log | tee -a $OUT/summary.txt
echo "== head A/B (FMA vs tcgen05 tf32x3)" | tee -a $OUT/summary.txt
timeout 200 python tools/time_head.py 2>&1 | tail -5 | tee -a $OUT/summary.txt
echo "== bridge" | tee -a $OUT/summary.txt
timeout 300 python tools/time_bridge.py 2>&1 | tail -12 | tee -a $OUT/summary.txt
echo "== operator lines" | tee -a $OUT/summary.txt
timeout 300 python tools/time_ops.py 2>&1 | tail -8 | tee -a $OUT/summary.txt
echo "== ncu: tensor pipe of the two tcgen05 kernels" | tee -a $OUT/summary.txt
timeout 300 ncu --set full --clock-control none -k regex:ground_p1_tc -c 2 -o $OUT/prof_p1tc python tools/time_head.py cfg2 > $OUT/ncu_p1tc.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:ground_fwd_kernel -c 2 -o $OUT/prof_fwd_fma python tools/time_head.py cfg2 > $OUT/ncu_fwd.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:gemm_bf16 -s 2 -c 2 -o $OUT/prof_gemm python tools/time_bridge.py > $OUT/ncu_gemm.log 2>&1
echo "== CPU lines" | tee -a $OUT/summary.txt
timeout 400 python tools/cpu_lines.py 2 2>&1 | tail -6 | tee -a $OUT/summary.txt
ls $OUT | tee -a $OUT/summary.txt
