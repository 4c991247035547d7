#!/usr/bin/env bash
# First GPU call after `git apply scripts/pending/lean_issue_loops.patch && python -m pytensor_federated_b200.build --force`:
# kernel tests, then the benches the patch is expected to move.  Everything under `timeout`.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > $OUT/pytest_lean.log 2>&1
rc=$?; echo "pytest exit $rc" >> $OUT/pytest_lean.log; tail -5 $OUT/pytest_lean.log
[[ $rc != 0 ]] && { echo "kernel tests failed: revert the patch (git apply -R) and bisect per kernel"; exit 1; }
for args in "--kernel tc" "--kernel fp8" "--kernel tc --chains 8" "--kernel tc --chains 16"; do
  tag=$(echo $args | tr -d '-' | tr ' ' '_')
  timeout 300 python bench.py $args --steps 20 --warmup 3 --out $OUT/bench_lean.jsonl > $OUT/bench_lean_$tag.log 2>&1
  echo "$args -> $(tail -1 $OUT/bench_lean_$tag.log | python -c 'import sys,json; d=json.loads(sys.stdin.readline()); print(d["value"], d["unit"], d["ms_per_step"], "ms")' 2>/dev/null || tail -2 $OUT/bench_lean_$tag.log)"
done
echo "before the patch: tc 173-177, fp8 297, chains 8: 1184, chains 16: 1694 (chain-)evals/s"
