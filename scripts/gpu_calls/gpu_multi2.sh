#!/usr/bin/env bash
# 2-GPU validation with tight timeouts (every minute on this box costs 2 GPU-minutes).
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 400 python -m pytest tests/test_multigpu.py -m gpu -q 2>&1 | tail -4 | tee $OUT/pytest_multigpu2.log
for B in fused collective; do
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2956$((RANDOM%10)) \
      benchmarks/bench_linreg_multi.py --backend $B --evals 3000 --out $OUT/linreg_multi.jsonl 2>&1 | grep -E "^\{" | cut -c1-400
done
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 \
    bench.py --impl nccl --gpus 2 --steps 10 --warmup 3 --out $OUT/bench_nccl_n2.jsonl 2>&1 | grep -E "^\{" | cut -c1-300
echo "nccl exit $?"
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29575 \
    bench.py --gpus 2 --steps 30 --warmup 5 --out $OUT/bench_n2_v2.jsonl 2>&1 | grep -E "^\{" | cut -c1-300
echo done
