#!/usr/bin/env bash
# Round-2 (8 GPUs): every BASELINE.json config that names 8 GPUs, the multi-GPU tests at world 8, the
# per-CTA phase breakdown and the NCCL baseline ("baseline B").  Everything under `timeout`.
set -u
OUT=gpurun_out; mkdir -p $OUT
run8() { # name, args...
  local name=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $PORT "$@" > $OUT/$name.log 2>&1
  echo "== $name rc=$?"; tail -1 $OUT/$name.log | cut -c1-420
  PORT=$((PORT+1))
}
PORT=29600
run8 bench_n8_glm bench.py --gpus 8 --steps 30 --warmup 5 --out $OUT/configs_r2.jsonl
run8 trace_n8 benchmarks/trace_breakdown.py --shards 8 --out $OUT/trace_n8.jsonl
run8 bench_n8_fp8 bench.py --gpus 8 --config fp8 --steps 30 --warmup 5 --out $OUT/configs_r2.jsonl
run8 bench_n8_linreg bench.py --gpus 8 --config linreg --steps 200 --warmup 20 --out $OUT/configs_r2.jsonl
run8 bench_n8_nccl bench.py --gpus 8 --impl nccl --steps 20 --warmup 3 --out $OUT/configs_r2.jsonl
timeout 420 python -m pytest tests/test_multigpu.py -m gpu -x -q -k "auto-8 or tc-8 or tc-4-8 or fp8-8 or hanging and 8" > $OUT/pytest_n8.log 2>&1; tail -4 $OUT/pytest_n8.log
nvidia-smi topo -m > $OUT/topo_8gpu_r2.txt 2>&1
