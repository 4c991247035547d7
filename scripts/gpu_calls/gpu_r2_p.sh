#!/usr/bin/env bash
# 2 GPUs: early TMA loads A/B on the same box (phase breakdown, bf16 and fp8), then the N = 2 bench
set -u
OUT=gpurun_out; mkdir -p $OUT
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port "$@"; }
port=29600
for rep in 1 2; do
for off in 0 1; do
  export B200FED_NO_EARLY_LOADS=$off
  port=$((port+1)); run $port benchmarks/trace_breakdown.py --shards 2 --out $OUT/trace_p_off$off.jsonl > $OUT/trace_p.log 2>&1
  echo "early_loads_off=$off bf16: $(tail -1 $OUT/trace_p.log | cut -c1-120)"
  port=$((port+1)); run $port benchmarks/trace_breakdown.py --shards 2 --kernel fp8 --out $OUT/trace_p_off$off.jsonl > $OUT/trace_p.log 2>&1
  echo "early_loads_off=$off fp8: $(tail -1 $OUT/trace_p.log | cut -c1-120)"
done
done
unset B200FED_NO_EARLY_LOADS
port=$((port+1)); run $port bench.py --gpus 2 --steps 30 --warmup 5 --out $OUT/bench_p.jsonl > $OUT/bench_p.log 2>&1; tail -1 $OUT/bench_p.log | cut -c1-300
