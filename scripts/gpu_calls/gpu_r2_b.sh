#!/usr/bin/env bash
# Round-2 call B (1 GPU): validate prologue/epilogue/PDL changes, re-measure the phase breakdown.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > $OUT/pytest_b.log 2>&1; tail -3 $OUT/pytest_b.log
timeout 200 python benchmarks/trace_breakdown.py --shards 1 --out $OUT/trace_r2b.jsonl > $OUT/trace_b_s1.log 2>&1; tail -1 $OUT/trace_b_s1.log
B200FED_NO_PDL=1 timeout 200 python benchmarks/trace_breakdown.py --shards 1 --out $OUT/trace_r2b.jsonl > $OUT/trace_b_s1_nopdl.log 2>&1; tail -1 $OUT/trace_b_s1_nopdl.log
B200FED_NO_LL=1 timeout 200 python benchmarks/trace_breakdown.py --shards 1 --out $OUT/trace_r2b.jsonl > $OUT/trace_b_s1_noll.log 2>&1; tail -1 $OUT/trace_b_s1_noll.log
timeout 300 python bench.py --kernel tc --steps 20 --warmup 3 --out $OUT/bench_b.jsonl > $OUT/bench_b_tc.log 2>&1; tail -1 $OUT/bench_b_tc.log | cut -c1-300
timeout 300 python bench.py --kernel tc --shards 1 --steps 50 --warmup 3 --out $OUT/bench_b.jsonl > $OUT/bench_b_tc1.log 2>&1; tail -1 $OUT/bench_b_tc1.log | cut -c1-300
