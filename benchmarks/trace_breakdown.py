"""Where does the fixed per-evaluation cost of the fused GLM kernel go?

Runs the headline config's per-GPU work (``--shards`` x ``--rows`` x ``--features`` spread over the
ranks), enables the per-CTA phase stamps (``fed::stamp``), launches a few back-to-back evaluations
and prints one JSON line with the phase breakdown of the last one (all times in microseconds on the
``%globaltimer`` clock of each GPU, relative to the first CTA entry on that GPU).

    python benchmarks/trace_breakdown.py --shards 1                      # one GPU's share of N = 8
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        benchmarks/trace_breakdown.py --shards 8
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def summarize(tr: np.ndarray) -> dict:
    tr = tr.astype(np.float64)
    t0 = tr[:, 0].min()
    us = lambda a: {"min": round(float(a.min()) / 1e3, 2), "med": round(float(np.median(a)) / 1e3, 2),
                    "max": round(float(a.max()) / 1e3, 2)}
    out = {"ctas": int(tr.shape[0])}
    out["entry_skew"] = us(tr[:, 0] - t0)
    out["theta_acquired"] = us(tr[:, 1] - t0)
    if tr[:, 2].any():
        out["setup"] = us(tr[:, 2] - tr[:, 1])
        out["first_tile_after_setup"] = us(tr[:, 3] - tr[:, 2])
        out["last_load_issued"] = us(tr[:, 4] - t0)
        out["stream_time"] = us(tr[:, 4] - tr[:, 3])
        out["loop_done"] = us(tr[:, 5] - t0)
        out["drain_after_last_load"] = us(tr[:, 5] - tr[:, 4])
        out["partial_store"] = us(tr[:, 6] - tr[:, 5])
    out["exit"] = us(tr[:, 7] - t0)
    last = int(np.argmax(tr[:, 7]))
    out["tail_after_slowest_partial"] = round(float(tr[:, 7].max() - tr[:, 6].max()) / 1e3, 2) if tr[:, 6].any() else None
    out["kernel_total"] = round(float(tr[:, 7].max() - t0) / 1e3, 2)
    out["last_cta"] = last
    return out


def main(argv=None) -> None:
    p = argparse.ArgumentParser()
    p.add_argument("--shards", type=int, default=1)
    p.add_argument("--rows", type=int, default=10_000_000)
    p.add_argument("--features", type=int, default=256)
    p.add_argument("--kernel", default="tc")
    p.add_argument("--chains", type=int, default=1)
    p.add_argument("--launches", type=int, default=6)
    p.add_argument("--out", default=None)
    a = p.parse_args(argv)

    import torch
    import torch.distributed as dist

    from pytensor_federated_b200.models import Fp8GlmShards, GlmShards, synth_logistic_shard, synth_logistic_shard_fp8
    from pytensor_federated_b200.parallel import FederatedEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    own_pg = world > 1 and not dist.is_initialized()
    if own_pg:
        dist.init_process_group("nccl", device_id=dev)
    mine = [s for s in range(a.shards) if s % world == rank]
    Xs, ys, scs = [], [], []
    for s in mine:
        if a.kernel == "fp8":
            X, sc, y = synth_logistic_shard_fp8(a.rows, a.features, seed=1000 + s, device=dev)
            scs.append(sc)
        else:
            X, y, _ = synth_logistic_shard(a.rows, a.features, seed=1000 + s, device=dev)
        Xs.append(X)
        ys.append(y)
    if a.kernel == "fp8":
        model = Fp8GlmShards(Xs, scs, ys, groups=mine, n_groups=a.shards, n_chains=a.chains)
        n_groups = a.shards
    else:
        model = GlmShards(Xs, ys, n_groups=1, family="logistic", n_chains=a.chains, kernel=a.kernel)
        n_groups = 1
    eng = FederatedEngine(model, backend="fused", timeout=60.0)
    eng.enable_cta_trace(True)
    rng = np.random.default_rng(3)
    K, P = a.chains, a.features
    ic = rng.normal(size=(K, n_groups) if K > 1 else (n_groups,)).astype(np.float32) * 0.1
    beta = rng.normal(size=(K, P) if K > 1 else (P,)).astype(np.float32) * 0.02
    n = a.launches
    res = {}
    if rank == 0:
        for _ in range(3):
            eng.evaluate(ic, beta)
        eng.set_device_theta((ic, beta), enable=True)
        stream = eng.torch_stream()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        last = 0
        for _ in range(n):
            last = eng.launch()
        ev1.record(stream)
        eng.wait(last)
        ev1.synchronize()
        res["ms_per_eval_back_to_back"] = ev0.elapsed_time(ev1) / n
        prev = eng.trace(last - 1)
        cur = eng.trace(last)
        tr = eng.cta_trace()
        t0 = float(tr[:, 0].astype(np.float64).min())
        res["gap_prev_result_to_entry_us"] = round((t0 - prev[2]) / 1e3, 2)
        res["theta_released_us"] = round((cur[0] - t0) / 1e3, 2)
        res["node_partial_released_us"] = round((cur[1] - t0) / 1e3, 2)
        res["result_released_us"] = round((cur[2] - t0) / 1e3, 2)
        res["rank0"] = summarize(tr)
    else:
        eng.serve(max_epochs=3 + n)
        res[f"rank{rank}"] = summarize(eng.cta_trace())
    if world > 1:
        allres = [None] * world
        dist.all_gather_object(allres, res)
        if rank == 0:
            for r in allres[1:]:
                res.update(r)
    eng.shutdown()
    if world > 1:
        dist.barrier()
        if own_pg:
            dist.destroy_process_group()
    if rank == 0:
        res["config"] = {"world": world, "shards": a.shards, "rows": a.rows, "features": a.features, "kernel": a.kernel,
                         "chains": a.chains}
        line = json.dumps(res)
        print(line, flush=True)
        if a.out:
            with open(a.out, "a") as fh:
                fh.write(line + "\n")


if __name__ == "__main__":
    main()
