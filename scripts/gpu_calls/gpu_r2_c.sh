#!/usr/bin/env bash
# Round-2 call C (1 GPU): dynamic chunk scheduler + two-level double-double reduction.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > $OUT/pytest_c.log 2>&1; tail -15 $OUT/pytest_c.log
timeout 200 python benchmarks/trace_breakdown.py --shards 1 --out $OUT/trace_r2c.jsonl > $OUT/trace_c_s1.log 2>&1; tail -1 $OUT/trace_c_s1.log
timeout 200 python benchmarks/trace_breakdown.py --shards 8 --out $OUT/trace_r2c.jsonl > $OUT/trace_c_s8.log 2>&1; tail -1 $OUT/trace_c_s8.log
timeout 300 python bench.py --kernel tc --steps 20 --warmup 3 --out $OUT/bench_c.jsonl > $OUT/bench_c_tc.log 2>&1; tail -1 $OUT/bench_c_tc.log | cut -c1-300
timeout 300 python bench.py --kernel tc --shards 1 --steps 50 --warmup 3 --out $OUT/bench_c.jsonl > $OUT/bench_c_tc1.log 2>&1; tail -1 $OUT/bench_c_tc1.log | cut -c1-300
timeout 300 python bench.py --kernel tc --chains 16 --steps 20 --warmup 3 --out $OUT/bench_c.jsonl > $OUT/bench_c_tc16.log 2>&1; tail -1 $OUT/bench_c_tc16.log | cut -c1-300
timeout 300 python bench.py --kernel fp8 --steps 20 --warmup 3 --out $OUT/bench_c.jsonl > $OUT/bench_c_fp8.log 2>&1; tail -1 $OUT/bench_c_fp8.log | cut -c1-300
