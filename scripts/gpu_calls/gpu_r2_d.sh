#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > $OUT/pytest_d.log 2>&1; tail -5 $OUT/pytest_d.log
for k in 1 4 8 16; do
timeout 300 python bench.py --kernel tc --chains $k --steps 20 --warmup 3 --out $OUT/bench_d.jsonl > $OUT/bench_d_tc$k.log 2>&1; tail -1 $OUT/bench_d_tc$k.log | cut -c1-200
done
timeout 200 python benchmarks/trace_breakdown.py --shards 1 --chains 16 --out $OUT/trace_r2d.jsonl > $OUT/trace_d_c16.log 2>&1; tail -1 $OUT/trace_d_c16.log
