#!/usr/bin/env bash
# per-node output blocks in the SIMT / general-shape GLM kernels: full single-GPU suite + sanitizer on the new tests
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_q.log 2>&1; tail -6 $OUT/pytest_q.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "cuda_core_kernels_keep_per_node" > $OUT/memcheck_q.log 2>&1; grep -E "ERROR SUMMARY|passed|failed" $OUT/memcheck_q.log | tail -3
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "cuda_core_kernels_keep_per_node" > $OUT/racecheck_q.log 2>&1; grep -E "RACECHECK SUMMARY|passed|failed" $OUT/racecheck_q.log | tail -3
