#!/usr/bin/env bash
# Multi-GPU validation under `gpurun --gpus N`: tests, then the bench at N ranks.
set -u
OUT=gpurun_out
mkdir -p $OUT
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus_multi.csv 2>&1
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 1200 python -m pytest tests/test_multigpu.py -m gpu -q -s > $OUT/pytest_multigpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_multigpu.log
tail -25 $OUT/pytest_multigpu.log
for K in ${KERNELS:-auto}; do
  NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 30 --warmup 5 --kernel $K --out $OUT/bench_n${N}_$K.jsonl > $OUT/bench_n${N}_$K.log 2>&1
  echo "bench exit $?" >> $OUT/bench_n${N}_$K.log
  tail -3 $OUT/bench_n${N}_$K.log | cut -c1-1500
done
echo "gpu_multi done"
if [[ ${REFERENCE:-0} == 1 ]]; then
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
      bench.py --impl reference --gpus $N --steps 10 --warmup 3 --out $OUT/bench_ref_n${N}.jsonl > $OUT/bench_ref_n${N}.log 2>&1
  tail -2 $OUT/bench_ref_n${N}.log | cut -c1-1200
  timeout 1200 python bench.py --impl reference --gpus 1 --steps 10 --warmup 3 --out $OUT/bench_ref_n1.jsonl > $OUT/bench_ref_n1.log 2>&1
  tail -2 $OUT/bench_ref_n1.log | cut -c1-1200
fi
