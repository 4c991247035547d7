"""Lotka–Volterra parameter-estimation shards (forward sensitivities, RK4).

Workload from ``/root/repo/BASELINE.json`` ("federated ODE parameter estimation
([timepoints, theta] -> trajectory -> LL), 4 shards on 4 GPUs"); the reference describes the
use-case only in prose (``/root/reference/README.md:39-52``).  ``theta = (alpha, beta, gamma,
delta)``; every series has its own known initial state; Gaussian observation noise.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import numpy as np

from .base import ShardModel


def lv_rhs(u, v, th):
    a, b, g, d = th
    return a * u - b * u * v, d * u * v - g * v


class OdeShards(ShardModel):
    """Lotka–Volterra parameter estimation: every shard holds many observed time series; one evaluation
    integrates them (RK4 with forward sensitivities, ``csrc/ode.cu``) and returns the Gaussian
    log-likelihood and its gradient w.r.t. ``theta = (alpha, beta, gamma, delta)``.  The reference only
    describes this workload in prose (``/root/reference/README.md:39-52``)."""

    def __init__(self, ts: Sequence, y0s: Sequence, y_obs: Sequence, sigmas: Sequence[float], substeps: int = 8):
        import torch

        self.ts = [t.to(torch.float32).contiguous() for t in ts]            # [n_t]
        self.y0s = [y.to(torch.float32).contiguous() for y in y0s]          # [2, n_series]
        self.y_obs = [y.to(torch.float32).contiguous() for y in y_obs]      # [n_t, 2, n_series]
        self.sigmas = [float(s) for s in sigmas]
        self.substeps = int(substeps)
        self.device = self.ts[0].device
        self.n_inputs = 1
        self.n_theta_words = 4
        self.n_vals = 5

    def pack_theta(self, inputs, out: np.ndarray):
        (theta,) = inputs
        out.view(np.float32)[:4] = np.asarray(theta, dtype=np.float32).reshape(4)
        return None

    def unpack_result(self, vals: np.ndarray, ctx=None) -> List[np.ndarray]:
        v = np.asarray(vals, dtype=np.float64)
        return [np.asarray(v[0]), v[1:5].copy()]

    def attach(self, lib, handle) -> None:
        from ..ops import native

        n = len(self.ts)
        native.check(
            lib.b200_engine_set_ode(
                handle, n,
                native.void_p_array([t.data_ptr() for t in self.ts]),
                native.void_p_array([t.data_ptr() for t in self.y0s]),
                native.void_p_array([t.data_ptr() for t in self.y_obs]),
                (C.c_int * n)(*[y.shape[1] for y in self.y0s]),
                (C.c_int * n)(*[t.numel() for t in self.ts]),
                (C.c_float * n)(*self.sigmas),
                (C.c_int * n)(*([self.substeps] * n)),
            ),
            "set_ode",
        )

    # -- eager oracle: same RK4 discretisation, autograd for the gradient, float64 -----------
    def reference_partial(self, inputs) -> np.ndarray:
        import torch

        (theta,) = inputs
        th = torch.tensor(np.asarray(theta, dtype=np.float64).reshape(4), device=self.device, requires_grad=True)
        total = torch.zeros((), dtype=torch.float64, device=self.device)
        for t, y0, yo, sigma in zip(self.ts, self.y0s, self.y_obs, self.sigmas):
            u, v = y0[0].double(), y0[1].double()
            t_prev = 0.0
            for j in range(t.numel()):
                h = (float(t[j]) - t_prev) / self.substeps
                for _ in range(self.substeps):
                    k1 = lv_rhs(u, v, th)
                    k2 = lv_rhs(u + 0.5 * h * k1[0], v + 0.5 * h * k1[1], th)
                    k3 = lv_rhs(u + 0.5 * h * k2[0], v + 0.5 * h * k2[1], th)
                    k4 = lv_rhs(u + h * k3[0], v + h * k3[1], th)
                    u = u + h / 6 * (k1[0] + 2 * k2[0] + 2 * k3[0] + k4[0])
                    v = v + h / 6 * (k1[1] + 2 * k2[1] + 2 * k3[1] + k4[1])
                t_prev = float(t[j])
                ru = yo[j, 0].double() - u
                rv = yo[j, 1].double() - v
                total = total + (-0.5 * (ru * ru + rv * rv) / sigma**2).sum() - 2 * u.numel() * (
                    np.log(sigma) + 0.918938533204672742
                )
        (grad,) = torch.autograd.grad(total, th)
        return np.concatenate([[float(total.detach())], grad.cpu().numpy()])


def synth_lv_shard(n_series: int, n_t: int, *, seed: int, device, theta=(1.0, 0.4, 0.8, 0.2), sigma=0.1,
                   t_end: float = 6.0, substeps: int = 8):
    """Simulates noisy Lotka–Volterra observations with the same RK4 scheme."""
    import torch

    gen = torch.Generator(device="cpu")
    gen.manual_seed(seed)
    t = torch.linspace(t_end / n_t, t_end, n_t)
    y0 = torch.stack([1.0 + torch.rand(n_series, generator=gen), 0.5 + torch.rand(n_series, generator=gen)])
    u, v = y0[0].double(), y0[1].double()
    th = torch.tensor(theta, dtype=torch.float64)
    obs = torch.empty(n_t, 2, n_series)
    t_prev = 0.0
    for j in range(n_t):
        h = (float(t[j]) - t_prev) / substeps
        for _ in range(substeps):
            k1 = lv_rhs(u, v, th)
            k2 = lv_rhs(u + 0.5 * h * k1[0], v + 0.5 * h * k1[1], th)
            k3 = lv_rhs(u + 0.5 * h * k2[0], v + 0.5 * h * k2[1], th)
            k4 = lv_rhs(u + h * k3[0], v + h * k3[1], th)
            u = u + h / 6 * (k1[0] + 2 * k2[0] + 2 * k3[0] + k4[0])
            v = v + h / 6 * (k1[1] + 2 * k2[1] + 2 * k3[1] + k4[1])
        t_prev = float(t[j])
        obs[j, 0] = (u + sigma * torch.randn(n_series, generator=gen, dtype=torch.float64)).float()
        obs[j, 1] = (v + sigma * torch.randn(n_series, generator=gen, dtype=torch.float64)).float()
    return t.to(device), y0.to(device), obs.to(device), sigma
