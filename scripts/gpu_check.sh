#!/usr/bin/env bash
# Runs on the GPU box under gpurun: tests, smoke, bench, launch list, ncu capture.
# Everything that should come back is written under gpurun_out/.
set -u
OUT=gpurun_out
mkdir -p $OUT
STAGE=${1:-all}
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > $OUT/gpus.csv 2>&1

if [[ $STAGE == all || $STAGE == tests ]]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> $OUT/pytest_gpu.log
  tail -15 $OUT/pytest_gpu.log
  timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1
  echo "smoke exit $?" >> $OUT/smoke.log
  tail -8 $OUT/smoke.log
fi

if [[ $STAGE == all || $STAGE == bench ]]; then
  for K in ${KERNELS:-simt}; do
    timeout 600 python bench.py --rows 1000000 --steps 20 --warmup 3 --kernel $K > $OUT/bench_small_$K.log 2>&1
    tail -2 $OUT/bench_small_$K.log
    timeout 900 python bench.py --steps 30 --warmup 5 --kernel $K --out $OUT/bench_full_$K.jsonl > $OUT/bench_full_$K.log 2>&1
    tail -2 $OUT/bench_full_$K.log
  done
fi

if [[ $STAGE == all || $STAGE == ncu ]]; then
  for K in ${KERNELS:-simt}; do
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fed_ -c 60 --csv \
        --log-file $OUT/launches_$K.csv python bench.py --rows 2000000 --steps 3 --warmup 3 --kernel $K > $OUT/ncu_launches_$K.log 2>&1
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:fed_glm -s 3 -c 1 \
        -o $OUT/prof_glm_$K -f python bench.py --rows 2000000 --steps 2 --warmup 3 --kernel $K > $OUT/ncu_full_$K.log 2>&1
    tail -3 $OUT/ncu_full_$K.log
  done
fi
if [[ $STAGE == chains ]]; then
  for C in ${CHAINS:-4 8 16}; do
    timeout 900 python bench.py --steps 20 --warmup 3 --kernel tc --chains $C --out $OUT/bench_full_tc_chains.jsonl > $OUT/bench_full_tc_c$C.log 2>&1
    tail -1 $OUT/bench_full_tc_c$C.log
  done
fi
echo "gpu_check done"
