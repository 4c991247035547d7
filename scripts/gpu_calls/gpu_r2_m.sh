#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -k "ipc-4 or symm-4 or tc-4] or tc-4-4 or tc-16-4 or fp8-4" > $OUT/pytest_n4_more.log 2>&1; tail -5 $OUT/pytest_n4_more.log
