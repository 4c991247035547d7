"""Headline benchmark: logp+grad evals/sec of the 8-shard federated logistic GLM.

Config (BASELINE.json): 8 shards x 10M rows x 256 features, bf16 design matrix, one theta per
evaluation; shards are spread over the N GPUs (8/N per GPU => strong scaling).  Synthetic data,
random parameters.  Contract: see the driver's instructions (one JSON line on rank 0).

    python bench.py                       # N=1, default steps
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29511 bench.py --gpus 8 --steps 50 --warmup 5
    python bench.py --impl reference ...  # the reference arm (see baseline/)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "logp+grad evals/sec for 8-shard federated GLM"


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=30)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="b200", choices=["b200", "reference", "nccl"])
    p.add_argument("--rows", type=int, default=10_000_000, help="rows per shard")
    p.add_argument("--features", type=int, default=256)
    p.add_argument("--shards", type=int, default=8)
    p.add_argument("--kernel", default=os.environ.get("B200FED_GLM_KERNEL", "auto"), choices=["auto", "simt", "tc", "fp8"])
    p.add_argument("--chains", type=int, default=1)
    p.add_argument("--out", default=None, help="also append the JSON line to this file")
    return p.parse_args()


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region.

    NVML is polled from a thread every 5 ms (timed regions can be as short as tens of milliseconds
    at N = 8); falls back to an ``nvidia-smi -lms`` subprocess when pynvml is unavailable.
    """

    QUERY = (
        "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
        "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    )
    REASON_BITS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}

    def __init__(self, gpu_index: int = 0):
        self.gpu_index = gpu_index
        self.lines = []
        self.proc = None
        self.samples = []       # (sm_mhz, max_mhz, power_w, reason_mask)
        self._stop = threading.Event()
        self._nvml = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.gpu_index]) if visible and visible.split(",")[0].isdigit() else self.gpu_index
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return self
        except Exception:
            self._nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _poll(self):
        nv = self._nvml
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self._handle, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(self._handle, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(self._handle) / 1000.0
                self.samples.append((float(sm), float(mx), pw, int(get_reasons(self._handle))))
            except Exception:
                pass
            time.sleep(0.005)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self._nvml is not None:
            self._stop.set()
            self.thread.join(1.0)
            if not self.samples:
                return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["no NVML samples"]}
            reasons = set()
            for _, _, _, mask in self.samples:
                for bit, name in self.REASON_BITS.items():
                    if mask & bit:
                        reasons.add(name)
            return {
                "sm_mhz": statistics.median(s[0] for s in self.samples),
                "sm_max_mhz": max(s[1] for s in self.samples),
                "power_w_max": max(s[2] for s in self.samples),
                "samples": len(self.samples),
                "reasons": sorted(reasons),
                "source": "nvml",
            }
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(smax) if smax else None,
            "power_w_max": max(power) if power else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
            "source": "nvidia-smi",
        }


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from pytensor_federated_b200.models import Fp8GlmShards, GlmShards, synth_logistic_shard, synth_logistic_shard_fp8
    from pytensor_federated_b200.parallel import FederatedEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # ---- data: this rank's share of the 8 shards -------------------------------------------
    my_shards = [s for s in range(args.shards) if s % world == rank]
    Xs, ys, scs = [], [], []
    for s in my_shards:
        if args.kernel == "fp8":
            X, sc, y = synth_logistic_shard_fp8(args.rows, args.features, seed=1000 + s, device=dev)
            scs.append(sc)
        else:
            X, y, _ = synth_logistic_shard(args.rows, args.features, seed=1000 + s, device=dev)
        Xs.append(X)
        ys.append(y)
    torch.cuda.synchronize()
    backend = "fused" if args.impl == "b200" else "collective"
    n_groups = 1
    if args.kernel == "fp8":  # hierarchical config: one partial-pooling group (intercept) per shard
        n_groups = args.shards
        model = Fp8GlmShards(Xs, scs, ys, groups=my_shards, n_groups=n_groups, n_chains=args.chains)
    else:
        model = GlmShards(Xs, ys, n_groups=1, family="logistic", n_chains=args.chains, kernel=args.kernel)
    eng = FederatedEngine(model, backend=backend, timeout=120.0)

    rng = np.random.default_rng(7)
    P, K = args.features, args.chains

    def draw_theta():
        ic = rng.normal(size=(K, n_groups) if K > 1 else (n_groups,)).astype(np.float32) * 0.1
        beta = rng.normal(size=(K, P) if K > 1 else (P,)).astype(np.float32) * 0.02
        return ic, beta

    thetas = [draw_theta() for _ in range(args.steps + args.warmup)]
    W, S = args.warmup, args.steps

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    result = {}
    sampler = None
    if rank == 0:
        # ---- warm-up through the public API -------------------------------------------------
        for i in range(W):
            eng.evaluate(*thetas[i])
        barrier()
        sampler = ClockSampler(local_rank).start()
        # ---- device-timed region: K back-to-back fused evaluations --------------------------
        if backend == "fused":
            stream = eng.torch_stream()
            eng.set_device_theta(thetas[W], enable=True)
            launches0 = eng.kernel_launches
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            last = 0
            for _ in range(S):
                last = eng.launch()
            ev1.record(stream)
            eng.wait(last)
            ev1.synchronize()
            dev_ms = ev0.elapsed_time(ev1)
            launches = eng.kernel_launches - launches0
            eng.set_device_theta(thetas[W], enable=False)
        else:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for i in range(S):
                eng.evaluate_raw(thetas[W + i])
            ev1.record()
            ev1.synchronize()
            dev_ms = ev0.elapsed_time(ev1)
            launches = 0
        barrier()
        # ---- end-to-end region: public API, pinned H2D of theta + D2H of the result each step
        t0 = time.perf_counter()
        checksum = 0.0
        for i in range(S):
            logp, d_ic, d_beta = eng.evaluate(*thetas[W + i])
            checksum += float(np.sum(logp))
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        barrier()
        clocks = sampler.stop()
        result = dict(dev_ms=dev_ms, e2e_s=e2e_s, launches=launches, checksum=checksum, clocks=clocks)
    else:
        # peers: each phase serves exactly as many epochs as the root evaluates; their kernels wait
        # on the device for the root's epoch flag, the host only keeps the queue filled
        eng.serve(max_epochs=W)
        barrier()
        eng.serve(max_epochs=S)
        barrier()
        eng.serve(max_epochs=S)
        barrier()

    # max over ranks of the device-timed region (the root's kernels cannot finish before every
    # peer delivered its partial, so the root time already dominates; reduce anyway)
    if world > 1:
        t = torch.tensor([result.get("dev_ms", 0.0)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            result["dev_ms"] = float(t.item())
        peer_launches = torch.tensor([float(eng.kernel_launches if backend == "fused" else 0)], device=dev,
                                     dtype=torch.float64)
        dist.all_reduce(peer_launches, op=dist.ReduceOp.SUM)
    comm_mode = eng.comm_mode
    bytes_per_eval = model.bytes_per_eval()
    flops_per_eval = model.flops_per_eval()
    n_theta_words, n_vals = model.n_theta_words, model.n_vals
    eng.shutdown()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    ms_per_step = result["dev_ms"] / S
    value = 1000.0 / ms_per_step * K
    e2e_value = S / result["e2e_s"] * K
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    line = {
        "metric": METRIC,
        "value": value,
        "unit": "evals/s",
        "n_gpus": world,
        "steps": S,
        "warmup": W,
        "ms_per_step": ms_per_step,
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "fp8-e4m3 block-scaled (32x32 UE8M0)" if args.kernel == "fp8" else "bf16",
        "data": "synthetic",
        "impl": args.impl,
        "config": {
            "model": "hierarchical logistic GLM, one group intercept per shard (logp + gradient)" if args.kernel == "fp8"
            else "federated logistic GLM (logp + gradient)",
            "shards": args.shards,
            "rows_per_shard": args.rows,
            "features": args.features,
            "chains_per_eval": K,
            "global_batch": args.shards * args.rows,
            "seq_len": args.features,
            "parallelism": f"shard-parallel x{world} ({args.shards // world} shards/GPU)",
            "kernel": args.kernel,
            "backend": backend,
            "comm": comm_mode,
            "l2_policy": "inputs (>= 5 GB per GPU) are larger than the 126 MB L2; no flush needed",
            "hbm_bytes_per_eval_per_gpu": bytes_per_eval,
            "hbm_roofline_frac_of_measured": (bytes_per_eval / (ms_per_step * 1e-3)) / (hbm * 1e9),
            # three-term roofline of the fused broadcast->compute->reduce kernel (seconds per eval per GPU):
            # HBM stream of the shard, tensor-core time of the two skinny GEMMs (N padded to 16 columns each),
            # NVLink bytes (theta in, partial out) at the measured 770 GB/s peer bandwidth
            "roofline_s": {
                "hbm": bytes_per_eval / (hbm * 1e9),
                "tensor": flops_per_eval * (32.0 / max(1, 4 * K)) / (peaks.get("bf16_tflops", 1590.0) * 1e12),
                "nvlink": (n_theta_words * 4 + n_vals * 8) / 770e9,
                "bound": "hbm",
            },
        },
        "clocks": result["clocks"],
        "e2e": {
            "value": e2e_value,
            "unit": "evals/s",
            "h2d_bytes_per_step": n_theta_words * 4,
            "d2h_bytes_per_step": n_vals * 8 + 8,
        },
        "gpu_launches": int(result["launches"]) * world if backend == "fused" else 0,
        "checksum": result["checksum"],
    }
    out = json.dumps(line)
    print(out, flush=True)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "a") as fh:
            fh.write(out + "\n")


def run_reference(args):
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    try:
        import reference_arm
    except Exception as ex:  # noqa: BLE001
        print(json.dumps({"impl": "reference", "unavailable": f"reference arm not importable: {ex}"}))
        return
    reference_arm.main(args)


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
