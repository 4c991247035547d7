#!/usr/bin/env bash
# N=4 scaling point (product + reference arm), tight timeouts: every minute costs 4 GPU-minutes.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29581 \
    bench.py --gpus 4 --steps 40 --warmup 5 --out $OUT/bench_n4.jsonl 2>&1 | grep -E "^\{" | cut -c1-300
echo "product exit $?"
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29585 \
    bench.py --impl reference --gpus 4 --steps 10 --warmup 3 --out $OUT/bench_ref_n4.jsonl 2>&1 | grep -E "^\{" | cut -c1-300
echo "reference exit $?"
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29589 \
    benchmarks/bench_linreg_multi.py --backend fused --evals 3000 --out $OUT/linreg_multi_n4.jsonl 2>&1 | grep -E "^\{" | cut -c1-400
echo done
