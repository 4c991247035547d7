#!/usr/bin/env bash
# 2-GPU validation: multi-GPU tests, then the headline bench with and without LL words for GLM-sized results.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 400 python -m pytest tests/test_multigpu.py -m gpu -q -x > $OUT/pytest_multigpu_final.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_multigpu_final.log
tail -4 $OUT/pytest_multigpu_final.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --gpus 2 --steps 40 --warmup 5 --out $OUT/bench_n2_$name.jsonl > $OUT/bench_n2_$name.log 2>&1
  echo "bench $name exit $?"; tail -1 $OUT/bench_n2_$name.log | cut -c1-400
}
run default A=1
run llglm B200FED_LL_MAX_VALS=2048 B200FED_LL_MAX_THETA=2048
echo done
