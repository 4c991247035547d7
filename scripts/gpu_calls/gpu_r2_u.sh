#!/usr/bin/env bash
# 2 GPUs: multi-GPU tests incl. the speculative scenarios, then bench (glm, fp8, linreg) with both e2e modes
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q -k "spec" > $OUT/pytest_u_spec.log 2>&1; tail -6 $OUT/pytest_u_spec.log
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q > $OUT/pytest_u.log 2>&1; tail -4 $OUT/pytest_u.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29710 \
    benchmarks/run_configs.py --gpus 2 --out $OUT/configs_u_n2.jsonl glm fp8 linreg > $OUT/run_configs_u.log 2>&1
echo "rc=$?"; grep -E '^\{' $OUT/run_configs_u.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['metric'][:45], round(d['value'],1), d.get('verified'), d['e2e'])"
