#!/usr/bin/env bash
# Round-2 (2 GPUs): multi-GPU tests at world 2, bench + phase breakdown at N = 2.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q > $OUT/pytest_n2.log 2>&1; tail -5 $OUT/pytest_n2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 2 --steps 30 --warmup 5 --out $OUT/bench_n2.jsonl > $OUT/bench_n2.log 2>&1; tail -1 $OUT/bench_n2.log | cut -c1-400
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 \
    benchmarks/trace_breakdown.py --shards 2 --out $OUT/trace_n2.jsonl > $OUT/trace_n2.log 2>&1; tail -1 $OUT/trace_n2.log | cut -c1-1500
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 \
    bench.py --gpus 2 --config linreg --shards 2 --steps 200 --warmup 20 --out $OUT/bench_n2.jsonl > $OUT/bench_n2_linreg.log 2>&1; tail -1 $OUT/bench_n2_linreg.log | cut -c1-300
