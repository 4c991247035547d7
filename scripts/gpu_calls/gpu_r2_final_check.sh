#!/usr/bin/env bash
# last call of the round: both arms of the driver contract at N = 1 on the final binary, the full GPU suite, smoke
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python bench.py --impl reference --gpus 1 --steps 10 --warmup 3 > $OUT/final_ref_n1.log 2>&1; tail -1 $OUT/final_ref_n1.log | cut -c1-300
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > $OUT/final_b200_n1.log 2>&1; tail -1 $OUT/final_b200_n1.log | cut -c1-300
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/final_pytest.log 2>&1; tail -3 $OUT/final_pytest.log
timeout 300 python __graft_entry__.py --smoke > $OUT/final_smoke.log 2>&1; tail -5 $OUT/final_smoke.log
