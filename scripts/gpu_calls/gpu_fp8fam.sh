#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "fp8" > $OUT/pytest_fp8fam.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_fp8fam.log
tail -30 $OUT/pytest_fp8fam.log
timeout 200 python bench.py --kernel fp8 --steps 20 --warmup 3 --out $OUT/bench_fp8_v4.jsonl > $OUT/bench_fp8_v4.log 2>&1
tail -1 $OUT/bench_fp8_v4.log | cut -c1-300
