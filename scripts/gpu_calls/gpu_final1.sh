#!/usr/bin/env bash
# Driver-like 1-GPU validation: GPU suite, smoke, default bench.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu_final.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu_final.log
tail -4 $OUT/pytest_gpu_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke_final.log 2>&1
tail -2 $OUT/smoke_final.log
timeout 300 python bench.py > $OUT/bench_final_n1.json 2> $OUT/bench_final_n1.err
echo "bench exit $?"; tail -1 $OUT/bench_final_n1.json | cut -c1-600
