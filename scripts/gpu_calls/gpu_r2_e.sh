#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > $OUT/pytest_e.log 2>&1; tail -3 $OUT/pytest_e.log
for k in 16 8; do
timeout 300 python bench.py --kernel tc --chains $k --steps 10 --warmup 3 --out $OUT/bench_e.jsonl > $OUT/bench_e_tc$k.log 2>&1; tail -1 $OUT/bench_e_tc$k.log | cut -c1-200
done
timeout 300 python bench.py --out $OUT/bench_e.jsonl > $OUT/bench_e_default.log 2>&1; tail -1 $OUT/bench_e_default.log
for c in fp8 ode linreg; do
timeout 300 python bench.py --config $c --steps 20 --warmup 3 --out $OUT/bench_e.jsonl > $OUT/bench_e_$c.log 2>&1; tail -1 $OUT/bench_e_$c.log | cut -c1-600
done
