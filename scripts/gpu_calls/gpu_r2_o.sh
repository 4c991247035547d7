#!/usr/bin/env bash
# early TMA loads (first chunk claimed and loaded before theta arrives): tests, N = 1 bench, one-shard phase breakdown
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > $OUT/pytest_o.log 2>&1; tail -4 $OUT/pytest_o.log
timeout 200 python benchmarks/trace_breakdown.py --shards 1 --out $OUT/trace_o.jsonl > $OUT/trace_o.log 2>&1; tail -1 $OUT/trace_o.log | cut -c1-900
timeout 200 python benchmarks/trace_breakdown.py --shards 1 --kernel fp8 --out $OUT/trace_o.jsonl > $OUT/trace_o_fp8.log 2>&1; tail -1 $OUT/trace_o_fp8.log | cut -c1-900
timeout 300 python bench.py --steps 30 --warmup 5 --out $OUT/bench_o.jsonl > $OUT/bench_o_glm.log 2>&1; tail -1 $OUT/bench_o_glm.log | cut -c1-200
timeout 300 python bench.py --config fp8 --steps 20 --warmup 3 --out $OUT/bench_o.jsonl > $OUT/bench_o_fp8.log 2>&1; tail -1 $OUT/bench_o_fp8.log | cut -c1-200
