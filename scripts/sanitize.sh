#!/usr/bin/env bash
# compute-sanitizer passes over the fused kernels on tiny shapes (run under gpurun; slow by nature).
# The reference has no race detection at all; here memcheck / racecheck / synccheck
# are part of the GPU test matrix and their summaries are copied to profiles/.
set -u
OUT=gpurun_out
mkdir -p $OUT
cat > /tmp/sanitize_case.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from pytensor_federated_b200.models import GlmShards, LinregShards, Fp8GlmShards, make_demo_data, synth_logistic_shard
from pytensor_federated_b200.parallel import FederatedEngine
dev = torch.device("cuda:0")
x, y, s = make_demo_data()
with FederatedEngine(LinregShards([x, x], [y, y], [s, s], device=dev)) as e:
    print("linreg", e.evaluate(np.array(0.4), np.array(1.2))[0])
X, yy, _ = synth_logistic_shard(128 * 6 + 5, 256, seed=0, device=dev)
beta = (np.random.default_rng(0).normal(size=256) * 0.02).astype(np.float32)
for kernel in ("simt", "tc"):
    with FederatedEngine(GlmShards([X], [yy], kernel=kernel), grid=4) as e:
        print(kernel, e.evaluate(np.array([0.1]), beta)[0])
m = Fp8GlmShards.from_dense([X.float()], [yy])
with FederatedEngine(m, grid=4) as e:
    print("fp8", e.evaluate(np.array([0.1]), beta)[0])
PY
for TOOL in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $TOOL --print-limit 20 python /tmp/sanitize_case.py > $OUT/sanitizer_$TOOL.log 2>&1
  echo "== $TOOL exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|linreg|simt|tc |fp8" $OUT/sanitizer_$TOOL.log | tail -8
done
