#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > $OUT/pytest_k.log 2>&1; tail -3 $OUT/pytest_k.log
timeout 300 python bench.py --config fp8 --steps 30 --warmup 5 --out $OUT/bench_k.jsonl > $OUT/bench_k_fp8.log 2>&1; tail -1 $OUT/bench_k_fp8.log | cut -c1-160
