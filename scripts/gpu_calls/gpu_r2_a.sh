#!/usr/bin/env bash
# Round-2 call A (1 GPU): per-CTA phase breakdown + ncu captures of the fp8 and 16-chain kernels.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 200 python benchmarks/trace_breakdown.py --shards 1 --out $OUT/trace_r2.jsonl > $OUT/trace_s1.log 2>&1; tail -1 $OUT/trace_s1.log
timeout 200 python benchmarks/trace_breakdown.py --shards 8 --out $OUT/trace_r2.jsonl > $OUT/trace_s8.log 2>&1; tail -1 $OUT/trace_s8.log
timeout 200 python benchmarks/trace_breakdown.py --shards 1 --kernel fp8 --out $OUT/trace_r2.jsonl > $OUT/trace_fp8.log 2>&1; tail -1 $OUT/trace_fp8.log
timeout 200 python benchmarks/trace_breakdown.py --shards 1 --chains 16 --out $OUT/trace_r2.jsonl > $OUT/trace_c16.log 2>&1; tail -1 $OUT/trace_c16.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fed_glm_fp8 -s 3 -c 1 -f -o $OUT/prof_fp8_r2 \
    python bench.py --kernel fp8 --shards 2 --steps 2 --warmup 3 > $OUT/ncu_fp8_r2.log 2>&1; tail -2 $OUT/ncu_fp8_r2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fed_glm_tc -s 3 -c 1 -f -o $OUT/prof_tc16_r2 \
    python bench.py --kernel tc --chains 16 --shards 2 --steps 2 --warmup 3 > $OUT/ncu_tc16_r2.log 2>&1; tail -2 $OUT/ncu_tc16_r2.log
ls -la $OUT/*.ncu-rep
