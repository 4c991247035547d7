#!/usr/bin/env bash
# compute-sanitizer passes over the fused kernels on small shapes (run under gpurun; slow by nature).
# The reference has no race detection at all; here memcheck / racecheck / synccheck
# are part of the GPU test matrix and their summaries are copied to profiles/.
#   bash scripts/sanitize.sh            # one GPU: every kernel, several chunks per CTA, chains, per-node blocks
#   bash scripts/sanitize.sh multi      # additionally (>= 2 GPUs): the cross-GPU protocols under racecheck / memcheck
set -u
OUT=gpurun_out
mkdir -p $OUT
cat > /tmp/sanitize_case.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from pytensor_federated_b200.models import (GlmShards, LinregShards, Fp8GlmShards, OdeShards, LOTKA_VOLTERRA, make_demo_data,
                                            synth_logistic_shard, synth_lv_shard)
from pytensor_federated_b200.parallel import FederatedEngine
dev = torch.device("cuda:0")
x, y, s = make_demo_data()
with FederatedEngine(LinregShards([x, x], [y, y], [s, s], device=dev)) as e:
    print("linreg", e.evaluate(np.array(0.4), np.array(1.2))[0])
# 47 + 12 tiles in two segments, 4 CTAs: every CTA claims several chunks, empty tiles pad the odd segment
Xa, ya, _ = synth_logistic_shard(128 * 46 + 5, 256, seed=0, device=dev)
Xb, yb, _ = synth_logistic_shard(128 * 12, 256, seed=1, device=dev)
beta = (np.random.default_rng(0).normal(size=256) * 0.02).astype(np.float32)
for kernel in ("simt", "tc"):
    with FederatedEngine(GlmShards([Xa, Xb], [ya, yb], kernel=kernel), grid=4) as e:
        print(kernel, e.evaluate(np.array([0.1]), beta)[0])
for K in (4, 16):
    b = (np.random.default_rng(K).normal(size=(K, 256)) * 0.02).astype(np.float32)
    with FederatedEngine(GlmShards([Xa, Xb], [ya, yb], kernel="tc", n_chains=K), grid=4) as e:
        print(f"tc-{K}", e.evaluate(np.zeros((K, 1)), b)[0][:2])
with FederatedEngine(GlmShards([Xa, Xb], [ya, yb], kernel="tc", node_ids=[0, 1], n_nodes=2), grid=4) as e:
    print("tc-nodes", e.evaluate(np.array([0.1]), beta)[0])
m = Fp8GlmShards.from_dense([Xa.float(), Xb.float()], [ya, yb])
with FederatedEngine(m, grid=4) as e:
    print("fp8", e.evaluate(np.array([0.1]), beta)[0])
sh = synth_lv_shard(64, 6, seed=0, device=dev)
for system in (None, LOTKA_VOLTERRA):
    with FederatedEngine(OdeShards([sh[0]], [sh[1]], [sh[2]], [sh[3]], system=system)) as e:
        print("ode" if system is None else "ode-generic", e.evaluate(np.array([1.0, 0.4, 0.8, 0.2]))[0])
PY
cat > /tmp/sanitize_multi.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch, torch.distributed as dist
from pytensor_federated_b200.models import GlmShards, LinregShards, make_demo_data, synth_logistic_shard
from pytensor_federated_b200.parallel import FederatedEngine
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
x, y, s = make_demo_data(seed=rank)
models = [LinregShards([x], [y], [s], local_ids=[rank], n_shards_total=world, device=dev)]
X, yy, _ = synth_logistic_shard(128 * 20 + 3, 256, seed=rank, device=dev)
models.append(GlmShards([X], [yy], kernel="tc"))
beta = (np.random.default_rng(0).normal(size=256) * 0.02).astype(np.float32)
for m in models:
    eng = FederatedEngine(m, timeout=120.0, comm=os.environ.get("SANITIZE_COMM", "ipc"))
    if rank == 0:
        for _ in range(3):
            out = eng.evaluate(np.array(0.4), np.array(1.2)) if isinstance(m, LinregShards) else eng.evaluate(np.array([0.1]), beta)
        print(type(m).__name__, float(np.sum(out[0])), eng.comm_mode, flush=True)
    else:
        eng.serve(max_epochs=3)
    eng.shutdown()
dist.barrier(); dist.destroy_process_group()
PY
for TOOL in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $TOOL --print-limit 20 python /tmp/sanitize_case.py > $OUT/sanitizer_$TOOL.log 2>&1
  echo "== $TOOL exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^linreg|^simt|^tc|^fp8|^ode" $OUT/sanitizer_$TOOL.log | tail -14
done
if [[ "${1:-}" == "multi" && $(nvidia-smi -L | wc -l) -ge 2 ]]; then
  for TOOL in racecheck memcheck; do
    timeout 1500 compute-sanitizer --tool $TOOL --target-processes all --print-limit 20 \
        python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29677 /tmp/sanitize_multi.py \
        > $OUT/sanitizer_multi_$TOOL.log 2>&1
    echo "== multi $TOOL exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^LinregShards|^GlmShards" $OUT/sanitizer_multi_$TOOL.log | tail -8
  done
fi
