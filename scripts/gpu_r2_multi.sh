#!/usr/bin/env bash
# usage: gpu_r2_multi.sh N "runs..." [pytest -k expr]
set -u
N=$1; RUNS=$2; KEXPR=${3:-}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29650 \
    benchmarks/run_configs.py --gpus $N --out $OUT/configs_r2_n$N.jsonl $RUNS > $OUT/run_configs_n$N.log 2>&1
echo "run_configs rc=$?"; grep -E '^\{|^run ' $OUT/run_configs_n$N.log | cut -c1-330
if [[ -n "$KEXPR" ]]; then
  timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q -k "$KEXPR" > $OUT/pytest_n$N.log 2>&1; tail -4 $OUT/pytest_n$N.log
fi
nvidia-smi topo -m > $OUT/topo_${N}gpu_r2.txt 2>&1
