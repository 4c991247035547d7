#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > $OUT/pytest_g.log 2>&1; tail -6 $OUT/pytest_g.log
timeout 300 python bench.py --config fp8 --steps 20 --warmup 3 --out $OUT/bench_g.jsonl > $OUT/bench_g_fp8.log 2>&1; tail -1 $OUT/bench_g_fp8.log | cut -c1-200
timeout 300 python bench.py --steps 30 --warmup 5 --out $OUT/bench_g.jsonl > $OUT/bench_g_glm.log 2>&1; tail -1 $OUT/bench_g_glm.log | cut -c1-200
timeout 200 python benchmarks/trace_breakdown.py --shards 1 --kernel fp8 --out $OUT/trace_g.jsonl > $OUT/trace_g_fp8.log 2>&1; tail -1 $OUT/trace_g_fp8.log | cut -c1-400
timeout 600 python benchmarks/bench_batching_gpu.py --chains 8 16 --evals 300 --out $OUT/batching_gpu_r2.jsonl > $OUT/batching_g.log 2>&1; tail -4 $OUT/batching_g.log | cut -c1-420
