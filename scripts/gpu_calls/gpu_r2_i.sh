#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > $OUT/pytest_i.log 2>&1; tail -5 $OUT/pytest_i.log
timeout 300 python bench.py --steps 30 --warmup 5 --out $OUT/bench_i.jsonl > $OUT/bench_i_glm.log 2>&1; tail -1 $OUT/bench_i_glm.log | cut -c1-160
timeout 300 python bench.py --features 192 --steps 20 --warmup 3 --out $OUT/bench_i.jsonl > $OUT/bench_i_p192.log 2>&1; tail -1 $OUT/bench_i_p192.log | cut -c1-160
bash scripts/sanitize.sh > $OUT/sanitize_i.log 2>&1; tail -40 $OUT/sanitize_i.log
