#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -k "ipc-8 or symm-8 or simt-8 or tc-4-8 or tc-16-8 or fp8-8" > $OUT/pytest_n8_more.log 2>&1; tail -5 $OUT/pytest_n8_more.log
