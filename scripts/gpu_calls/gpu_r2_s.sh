#!/usr/bin/env bash
# 8 GPUs, final binary: headline GLM + phase breakdown + fp8 + linreg in one process group
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 \
    benchmarks/run_configs.py --gpus 8 --out $OUT/configs_r2_final_n8.jsonl glm trace fp8 linreg > $OUT/run_configs_s.log 2>&1
echo "rc=$?"; grep -E '^\{' $OUT/run_configs_s.log | cut -c1-330
