"""Diagnostics for the tensor-core GLM kernels: prints error structure instead of just failing.

    python scripts/tc_debug.py bf16|fp8
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pytensor_federated_b200.models import Fp8GlmShards, GlmShards, synth_logistic_shard
from pytensor_federated_b200.parallel import FederatedEngine

dev = torch.device("cuda:0")
MODE = sys.argv[1] if len(sys.argv) > 1 else "bf16"


def report(tag, m, ic, beta):
    try:
        with FederatedEngine(m, timeout=10) as eng:
            got = eng.evaluate(ic, beta)
    except Exception as ex:  # noqa: BLE001
        print(f"{tag}: ERROR {ex}")
        return False
    want = m.unpack_result(m.reference_partial([ic, beta], dtype=torch.float64))
    g, w = got[2], want[2]
    corr = float(np.corrcoef(g, w)[0, 1]) if np.std(g) > 0 else float("nan")
    print(f"{tag}: logp {float(got[0]):.5f} vs {float(want[0]):.5f} | d_ic {got[1]} vs {want[1]} | "
          f"grad corr {corr:.6f} maxabs {np.abs(g - w).max():.4g} scale {np.abs(w).max():.4g}")
    if not np.allclose(g, w, rtol=1e-3, atol=1e-3 * np.abs(w).max()):
        print("   got ", np.round(g[:8], 4), np.round(g[-8:], 4))
        print("   want", np.round(w[:8], 4), np.round(w[-8:], 4))
        print("   ratio", np.round(g[:8] / w[:8], 4), np.round(g[-8:] / w[-8:], 4))
    return True


beta_rng = np.random.default_rng(2)
if MODE == "bf16":
    for P, rows, fam in [(256, [128], "gaussian"), (256, [128 * 3 + 5], "logistic"), (128, [1000], "logistic"),
                         (256, [20000, 333], "logistic")]:
        Xs, ys = [], []
        for i, n in enumerate(rows):
            X, y, _ = synth_logistic_shard(n, P, seed=i, device=dev)
            Xs.append(X)
            ys.append(y)
        report(f"bf16 P={P} rows={rows} {fam}", GlmShards(Xs, ys, family=fam, kernel="tc"), np.array([0.25]),
               (beta_rng.normal(size=P) * 0.03).astype(np.float32))
else:
    for P, rows, hetero in [(256, [128], 0), (256, [128], 1), (256, [128], 2), (256, [128 * 5 + 3], 2), (128, [2000, 300], 2)]:
        torch.manual_seed(1)
        Xs, ys = [], []
        for n in rows:
            X = torch.randn(n, P, device=dev)
            if hetero >= 1:  # feature-dependent magnitudes -> scales differ between feature blocks
                X = X * torch.exp(torch.randn(P, device=dev))
            if hetero >= 2:  # row-dependent magnitudes -> scales differ between row groups
                X = X * torch.exp(0.5 * torch.randn(n, 1, device=dev))
            Xs.append(X)
            ys.append((torch.rand(n, device=dev) < 0.4).float())
        ok = report(f"fp8 P={P} rows={rows} hetero={hetero}", Fp8GlmShards.from_dense(Xs, ys), np.array([0.25]),
                    (beta_rng.normal(size=P) * 0.02).astype(np.float32))
        if not ok:
            break  # a device fault poisons the context
