"""Tiny driver for ncu captures of the latency kernels (linreg, ODE, general-shape GLM)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pytensor_federated_b200.models import GlmShards, LinregShards, OdeShards, make_demo_data, synth_lv_shard
from pytensor_federated_b200.parallel import FederatedEngine

dev = torch.device("cuda:0")
x, y, s = make_demo_data()
with FederatedEngine(LinregShards([x] * 8, [y] * 8, [s] * 8, device=dev)) as e:
    for i in range(6):
        e.evaluate(np.array(0.1 * i), np.array(0.5))
sh = [synth_lv_shard(20_000, 32, seed=k, device=dev) for k in range(2)]
with FederatedEngine(OdeShards([a[0] for a in sh], [a[1] for a in sh], [a[2] for a in sh], [a[3] for a in sh])) as e:
    for i in range(6):
        e.evaluate(np.array([1.0, 0.4, 0.8, 0.2]))
X = torch.randn(400_000, 200, device=dev)
with FederatedEngine(GlmShards([X], [(torch.rand(400_000, device=dev) < 0.5).float()])) as e:
    for i in range(6):
        e.evaluate(np.array([0.1]), np.zeros(200, dtype=np.float32))
