#!/bin/bash
OUT=${1:-gpurun_out/iter7}
mkdir -p $OUT
echo "== pytest -m gpu (all, no -x)" | tee $OUT/summary.txt
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/summary.txt; tail -15 $OUT/pytest.log | tee -a $OUT/summary.txt
echo "== operator lines" | tee -a $OUT/summary.txt
timeout 300 python tools/time_ops.py 2>&1 | tail -10 | tee -a $OUT/summary.txt
timeout 300 python tools/time_ops.py --cfg cfg2_real 2>&1 | tail -10 | tee -a $OUT/summary.txt
