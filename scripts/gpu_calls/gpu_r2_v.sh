#!/usr/bin/env bash
# 8 GPUs, final binary with speculative launches in the e2e region: GLM, fp8, linreg in one process group,
# then the two speculative multi-GPU scenarios at world 8
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29720 \
    benchmarks/run_configs.py --gpus 8 --out $OUT/configs_v_n8.jsonl glm fp8 linreg > $OUT/run_configs_v.log 2>&1
echo "rc=$?"; grep -E '^\{' $OUT/run_configs_v.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['metric'][:45], round(d['value'],1), d.get('verified'), d['e2e'])"
timeout 300 python -m pytest tests/test_multigpu.py -m gpu -x -q -k "spec-8" > $OUT/pytest_v_spec.log 2>&1; tail -3 $OUT/pytest_v_spec.log
