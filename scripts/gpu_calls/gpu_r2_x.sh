#!/usr/bin/env bash
# 4 GPUs: federated linear regression under NUTS (one shard pair per GPU), after the graph-IR work
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29730 \
    bench.py --gpus 4 --config linreg --steps 200 --warmup 20 --out $OUT/bench_x_n4.jsonl > $OUT/bench_x.log 2>&1
echo "rc=$?"; tail -1 $OUT/bench_x.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'], d.get('nuts'))"
