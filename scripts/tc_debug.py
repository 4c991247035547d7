"""Diagnostics for the tcgen05 GLM kernel: prints error structure instead of just failing."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pytensor_federated_b200.models import GlmShards, synth_logistic_shard
from pytensor_federated_b200.parallel import FederatedEngine

dev = torch.device("cuda:0")
for P, rows, fam in [(256, [128], "gaussian"), (256, [128 * 3 + 5], "logistic"), (128, [1000], "logistic"), (256, [20000, 333], "logistic")]:
    Xs, ys = [], []
    for i, n in enumerate(rows):
        X, y, _ = synth_logistic_shard(n, P, seed=i, device=dev)
        Xs.append(X); ys.append(y)
    beta = (np.random.default_rng(2).normal(size=P) * 0.03).astype(np.float32)
    ic = np.array([0.25])
    m = GlmShards(Xs, ys, family=fam, kernel="tc")
    try:
        with FederatedEngine(m, timeout=10) as eng:
            got = eng.evaluate(ic, beta)
    except Exception as ex:
        print(f"P={P} rows={rows} {fam}: ERROR {ex}")
        continue
    want = m.unpack_result(m.reference_partial([ic, beta], dtype=torch.float64))
    g, w = got[2], want[2]
    corr = float(np.corrcoef(g, w)[0, 1]) if np.std(g) > 0 else float("nan")
    print(f"P={P} rows={rows} {fam}: logp {float(got[0]):.6f} vs {float(want[0]):.6f} | d_ic {got[1]} vs {want[1]} | "
          f"grad corr {corr:.6f} maxabs {np.abs(g - w).max():.4g} scale {np.abs(w).max():.4g}")
    if not np.allclose(g, w, rtol=1e-3, atol=1e-2 * np.abs(w).max()):
        print("   got ", np.round(g[:12], 3)); print("   want", np.round(w[:12], 3))
        # does a permutation explain it?
        order_g, order_w = np.argsort(g), np.argsort(w)
        print("   sorted-match:", np.allclose(np.sort(g), np.sort(w), rtol=1e-3, atol=1e-2 * np.abs(w).max()),
              " first perm idx:", order_w[:8], "->", order_g[:8])


# ---- block-scaled fp8 kernel --------------------------------------------------------------------
from pytensor_federated_b200.models import Fp8GlmShards
for P, rows, hetero in [(256, [128], False), (256, [128], True), (256, [128 * 5 + 3], True), (128, [2000, 300], True)]:
    torch.manual_seed(1)
    Xs, ys = [], []
    for n in rows:
        X = torch.randn(n, P, device=dev)
        if hetero:
            X = X * torch.exp(torch.randn(P, device=dev)) * torch.exp(0.5 * torch.randn(n, 1, device=dev))
        Xs.append(X); ys.append((torch.rand(n, device=dev) < 0.4).float())
    m = Fp8GlmShards.from_dense(Xs, ys)
    beta = (np.random.default_rng(2).normal(size=P) * 0.02).astype(np.float32)
    ic = np.array([0.25])
    try:
        with FederatedEngine(m, timeout=10) as eng:
            got = eng.evaluate(ic, beta)
    except Exception as ex:
        print(f"fp8 P={P} rows={rows} hetero={hetero}: ERROR {ex}")
        continue
    want = m.unpack_result(m.reference_partial([ic, beta], dtype=torch.float64))
    g, w = got[2], want[2]
    corr = float(np.corrcoef(g, w)[0, 1]) if np.std(g) > 0 else float("nan")
    print(f"fp8 P={P} rows={rows} hetero={hetero}: logp {float(got[0]):.5f} vs {float(want[0]):.5f} | d_ic {got[1]} vs {want[1]} | "
          f"grad corr {corr:.6f} maxabs {np.abs(g - w).max():.4g} scale {np.abs(w).max():.4g}")
    if not np.allclose(g, w, rtol=1e-3, atol=1e-3 * np.abs(w).max()):
        print("   got ", np.round(g[:8], 4), np.round(g[128:136], 4) if P > 128 else "")
        print("   want", np.round(w[:8], 4), np.round(w[128:136], 4) if P > 128 else "")
        print("   ratio", np.round(g[:8] / w[:8], 4))
