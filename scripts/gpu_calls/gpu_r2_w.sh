#!/usr/bin/env bash
# linreg under NUTS after the graph-IR work (N = 1 here; the caller adds --gpus for N = 2)
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python bench.py --config linreg --steps 200 --warmup 20 --out $OUT/bench_w.jsonl > $OUT/bench_w_linreg.log 2>&1; tail -1 $OUT/bench_w_linreg.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d.get('nuts'))"
timeout 300 python benchmarks/bench_configs.py --out $OUT/configs_misc_w.jsonl > $OUT/bench_configs_w.log 2>&1; tail -4 $OUT/bench_configs_w.log | cut -c1-400
