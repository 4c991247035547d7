#!/usr/bin/env bash
# fp8 kernel with the dynamic chunk scheduler: tests, bench, phase breakdown
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > $OUT/pytest_h.log 2>&1; tail -6 $OUT/pytest_h.log
timeout 300 python bench.py --config fp8 --steps 20 --warmup 3 --out $OUT/bench_h.jsonl > $OUT/bench_h_fp8.log 2>&1; tail -1 $OUT/bench_h_fp8.log | cut -c1-200
timeout 300 python bench.py --config fp8 --chains 3 --steps 20 --warmup 3 --out $OUT/bench_h.jsonl > $OUT/bench_h_fp8c3.log 2>&1; tail -1 $OUT/bench_h_fp8c3.log | cut -c1-200
timeout 200 python benchmarks/trace_breakdown.py --shards 1 --kernel fp8 --out $OUT/trace_h.jsonl > $OUT/trace_h_fp8.log 2>&1; tail -1 $OUT/trace_h_fp8.log | cut -c1-600
timeout 300 python bench.py --steps 30 --warmup 5 --out $OUT/bench_h.jsonl > $OUT/bench_h_glm.log 2>&1; tail -1 $OUT/bench_h_glm.log | cut -c1-200
timeout 300 python bench.py --chains 16 --steps 10 --warmup 3 --out $OUT/bench_h.jsonl > $OUT/bench_h_glm16.log 2>&1; tail -1 $OUT/bench_h_glm16.log | cut -c1-200
timeout 600 python benchmarks/bench_batching_gpu.py --chains 8 16 --evals 300 --out $OUT/batching_gpu_r2.jsonl > $OUT/batching_h.log 2>&1; tail -4 $OUT/batching_h.log | cut -c1-420
