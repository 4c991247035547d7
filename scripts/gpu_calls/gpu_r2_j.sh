#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > $OUT/pytest_j.log 2>&1; tail -3 $OUT/pytest_j.log
timeout 300 python bench.py --steps 30 --warmup 5 --out $OUT/bench_j.jsonl > $OUT/bench_j_glm.log 2>&1; tail -1 $OUT/bench_j_glm.log | cut -c1-160
timeout 300 python bench.py --config fp8 --steps 30 --warmup 5 --out $OUT/bench_j.jsonl > $OUT/bench_j_fp8.log 2>&1; tail -1 $OUT/bench_j_fp8.log | cut -c1-160
bash scripts/sanitize.sh multi > $OUT/sanitize_j.log 2>&1; tail -60 $OUT/sanitize_j.log
