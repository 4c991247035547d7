#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 100 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu_final.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu_final.log
tail -3 $OUT/pytest_gpu_final.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke_final.log 2>&1
tail -1 $OUT/smoke_final.log
