#!/bin/bash
OUT=${1:-gpurun_out/final1}
mkdir -p $OUT
echo "== pytest -m gpu" | tee $OUT/summary.txt
timeout 300 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/summary.txt; tail -4 $OUT/pytest.log | tee -a $OUT/summary.txt
timeout 120 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err
echo "bench rc=$?" | tee -a $OUT/summary.txt
cat $OUT/bench_default.json | tee -a $OUT/summary.txt
