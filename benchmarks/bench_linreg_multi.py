"""Config 2 of BASELINE.json: federated Bayesian linear regression, one 10-row shard per GPU.

    torchrun --nproc-per-node 8 benchmarks/bench_linreg_multi.py [--backend fused|collective]

Latency-bound: the figure of merit is evaluations/s of the client (rank 0) and the device-side
phase times from the %globaltimer trace (theta released -> last partial arrived -> result released).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist

    from pytensor_federated_b200.models import LinregShards, make_demo_data
    from pytensor_federated_b200.parallel import FederatedEngine

    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="fused")
    ap.add_argument("--evals", type=int, default=5000)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    x, y, sigma = make_demo_data(seed=123 + rank)
    model = LinregShards([x], [y], [sigma], local_ids=[rank], n_shards_total=world, device=dev)
    eng = FederatedEngine(model, backend=args.backend, timeout=60.0)
    n = args.evals
    if rank == 0:
        rng = np.random.default_rng(0)
        thetas = [(rng.normal(size=world), np.asarray(rng.normal())) for _ in range(n + 200)]
        for th in thetas[:200]:
            eng.evaluate(*th)
        t0 = time.perf_counter()
        for th in thetas[200:]:
            eng.evaluate(*th)
        dt = time.perf_counter() - t0
        line = {"config": f"federated linreg, {world} shards x 10 rows on {world} GPU(s)", "backend": args.backend,
                "comm": eng.comm_mode, "evals_per_s": n / dt, "latency_us": 1e6 * dt / n}
        if args.backend == "fused":
            phases = np.array([eng.trace(e) for e in range(eng.n_evals - 100, eng.n_evals)], dtype=np.float64)
            line["device_us_theta_to_result_median"] = float(np.median(phases[:, 2] - phases[:, 0]) / 1e3)
            line["device_us_theta_to_local_partial_median"] = float(np.median(phases[:, 1] - phases[:, 0]) / 1e3)
        print(json.dumps(line), flush=True)
        if args.out:
            with open(args.out, "a") as fh:
                fh.write(json.dumps(line) + "\n")
        eng.shutdown()
    else:
        eng.serve()
        eng.shutdown()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
