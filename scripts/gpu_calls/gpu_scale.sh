#!/usr/bin/env bash
# Scaling run on one multi-GPU box: product + reference arms at each N (no tests).
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo_scale.txt 2>&1
PORT=29540
for N in ${NS:-8 4}; do
  PORT=$((PORT+3))
  NCCL_DEBUG=${NCCL_DEBUG:-WARN} timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $N --steps 40 --warmup 5 --out $OUT/bench_scale.jsonl > $OUT/bench_scale_n$N.log 2>&1
  echo "N=$N exit $?"; tail -1 $OUT/bench_scale_n$N.log | cut -c1-300
  if [[ ${REFERENCE:-1} == 1 ]]; then
    PORT=$((PORT+3))
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
        bench.py --impl reference --gpus $N --steps 10 --warmup 3 --out $OUT/bench_scale_ref.jsonl > $OUT/bench_scale_ref_n$N.log 2>&1
    echo "ref N=$N exit $?"; tail -1 $OUT/bench_scale_ref_n$N.log | cut -c1-300
  fi
  if [[ ${NCCL_ARM:-0} == 1 ]]; then
    PORT=$((PORT+3))
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
        bench.py --impl nccl --gpus $N --steps 20 --warmup 3 --out $OUT/bench_scale_nccl.jsonl > $OUT/bench_scale_nccl_n$N.log 2>&1
    echo "nccl N=$N exit $?"; tail -1 $OUT/bench_scale_nccl_n$N.log | cut -c1-300
  fi
done
echo "gpu_scale done"
