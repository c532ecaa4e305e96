#!/bin/bash
OUT=${1:-gpurun_out/iter10}
mkdir -p $OUT
echo "== pytest roi_align" | tee $OUT/summary.txt
timeout 600 python -m pytest tests/test_roi_align.py -m gpu -q > $OUT/pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/summary.txt; tail -4 $OUT/pytest.log | tee -a $OUT/summary.txt
timeout 300 python tools/time_ops.py 2>&1 | grep roi_align | tee -a $OUT/summary.txt
timeout 300 python tools/time_ops.py --cfg cfg2_real 2>&1 | grep roi_align | tee -a $OUT/summary.txt
timeout 300 python tools/time_ops.py --cfg cfg4 2>&1 | grep roi_align | tee -a $OUT/summary.txt


