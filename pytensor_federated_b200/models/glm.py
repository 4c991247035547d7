"""Generalised linear model shards with a low-precision, HBM-resident design matrix.

Workloads from ``/root/repo/BASELINE.json`` (not present in the reference, whose only model is
the 10-point linear regression): *federated logistic GLM, 10M rows x 256 features per shard,
bf16 design matrix* and *hierarchical GLM, 8 partial-pooling groups, fp8 block-scaled design
matrix*.  ``theta = [intercept[G], beta[P]]`` (float32) per chain; a segment (shard) uses
``intercept[group]``, so G = 1 is the pooled GLM and G = #shards the partial-pooling one.
Result per chain: ``[LL, dLL/dintercept[G], dLL/dbeta[P]]`` (float64).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from .base import ShardModel

FAMILIES = {"logistic": 0, "poisson": 1, "gaussian": 2}


class GlmShards(ShardModel):
    """The GLM segments that live on ONE GPU.

    Parameters
    ----------
    Xs, ys
        Per-segment design matrices ``[n_rows, P]`` (bf16, row-major) and responses ``[n_rows]``
        (float32).
    groups
        Intercept index of every segment.
    n_groups
        Total number of intercepts G in the federation.
    n_chains
        Parameter vectors evaluated per call (K).  ``K > 1`` needs the tensor-core kernel.
    kernel
        ``"simt"`` (one chain, any P % 8 == 0 up to 512), ``"tc"`` (tcgen05 + TMA) or ``"auto"``.
    """

    def __init__(
        self,
        Xs: Sequence,
        ys: Sequence,
        *,
        groups: Optional[Sequence[int]] = None,
        n_groups: int = 1,
        family: str = "logistic",
        n_chains: int = 1,
        kernel: str = "auto",
        scales: Optional[Sequence] = None,
    ) -> None:
        import torch

        if len(Xs) != len(ys):
            raise ValueError("Xs and ys must have the same length")
        self.Xs = list(Xs)
        self.ys = [y.to(torch.float32).contiguous() for y in ys]
        self.scales = list(scales) if scales is not None else None
        self.groups = list(groups) if groups is not None else [0] * len(Xs)
        self.n_groups = int(n_groups)
        self.family = family
        self.n_chains = int(n_chains)
        self.kernel = kernel
        X0 = self.Xs[0]
        self.n_features = int(X0.shape[1])
        self.ld = int(X0.stride(0))
        for X in self.Xs:
            if X.dim() != 2 or X.shape[1] != self.n_features or X.stride(1) != 1 or X.stride(0) != self.ld:
                raise ValueError("all design matrices must be row-major [n, P] with one row stride")
        self.device = X0.device
        self.n_inputs = 2
        self.n_params = self.n_groups + self.n_features
        self.n_theta_words = self.n_chains * self.n_params
        self.n_vals = self.n_chains * (1 + self.n_params)

    @property
    def n_rows(self) -> int:
        return int(sum(X.shape[0] for X in self.Xs))

    # -- packing ---------------------------------------------------------------------------
    def pack_theta(self, inputs, out: np.ndarray) -> None:
        intercept, beta = inputs
        th = out.view(np.float32).reshape(self.n_chains, self.n_params)
        ic = np.asarray(intercept, dtype=np.float32)
        bt = np.asarray(beta, dtype=np.float32)
        self._batched = bt.ndim == 2
        self._icpt_shape = ic.shape
        th[:, : self.n_groups] = ic.reshape(self.n_chains, -1) if self._batched else ic.reshape(1, -1)
        th[:, self.n_groups :] = bt.reshape(self.n_chains, self.n_features)

    _batched = False
    _icpt_shape = ()

    def unpack_result(self, vals: np.ndarray) -> List[np.ndarray]:
        v = np.asarray(vals, dtype=np.float64).reshape(self.n_chains, 1 + self.n_params)
        G = self.n_groups
        if self._batched:
            return [v[:, 0].copy(), v[:, 1 : 1 + G].reshape((self.n_chains,) + tuple(self._icpt_shape[1:])).copy(),
                    v[:, 1 + G :].copy()]
        return [np.asarray(v[0, 0]), v[0, 1 : 1 + G].reshape(self._icpt_shape).copy(), v[0, 1 + G :].copy()]

    # -- native ----------------------------------------------------------------------------
    def use_tensor_cores(self) -> bool:
        if self.kernel == "tc":
            return True
        if self.kernel == "simt":
            return False
        # auto: the tcgen05 kernel wherever its shape constraints hold (it is faster even for one
        # chain: TMA streaming + the X tile reused from smem for both GEMMs)
        import torch

        tc_ok = (
            self.n_features % 128 == 0 and 128 <= self.n_features <= 384 and self.n_chains <= 8
            and self.Xs[0].dtype == torch.bfloat16 and len(self.Xs) <= 64
        )
        if not tc_ok and self.n_chains > 1:
            raise ValueError("multi-chain evaluation needs the tensor-core kernel (P % 128 == 0, P <= 384, K <= 8)")
        return tc_ok

    def attach(self, lib, handle) -> None:
        from ..ops import native

        n = len(self.Xs)
        Xp = native.void_p_array([X.data_ptr() for X in self.Xs])
        yp = native.void_p_array([y.data_ptr() for y in self.ys])
        sp = native.void_p_array([s.data_ptr() for s in self.scales]) if self.scales else None
        rows = (C.c_longlong * n)(*[X.shape[0] for X in self.Xs])
        grp = (C.c_int * n)(*self.groups)
        native.check(
            lib.b200_engine_set_glm(
                handle, n, Xp, yp, sp, rows, grp, self.n_features, self.ld, self.n_groups,
                self.n_chains, FAMILIES[self.family], int(self.use_tensor_cores()),
            ),
            "set_glm",
        )

    # -- eager oracle (also the compute step of the NCCL baseline) ---------------------------
    def reference_partial(self, inputs, *, dtype=None) -> np.ndarray:
        import torch

        dtype = dtype or torch.float32
        intercept, beta = inputs
        ic = torch.as_tensor(np.asarray(intercept, dtype=np.float64)).reshape(self.n_chains, -1)
        bt = torch.as_tensor(np.asarray(beta, dtype=np.float64)).reshape(self.n_chains, self.n_features)
        out = torch.zeros(self.n_chains, 1 + self.n_params, dtype=torch.float64)
        for X, y, g in zip(self.Xs, self.ys, self.groups):
            Xf = self._dequant(X).to(dtype)
            B = bt.to(self.device, dtype)                          # [K, P]
            eta = Xf @ B.T + ic[:, g].to(self.device, dtype)        # [n, K]
            yy = y.to(dtype).unsqueeze(1)
            if self.family == "logistic":
                ll = yy * eta - torch.nn.functional.softplus(eta)
                r = yy - torch.sigmoid(eta)
            elif self.family == "poisson":
                mu = torch.exp(eta)
                ll = yy * eta - mu
                r = yy - mu
            else:
                d = yy - eta
                ll = -0.5 * d * d - 0.918938533204672742
                r = d
            out[:, 0] += ll.double().sum(0).cpu()
            out[:, 1 + g] += r.double().sum(0).cpu()
            out[:, 1 + self.n_groups :] += (r.T @ Xf).double().cpu()
        return out.reshape(-1).numpy()

    def _dequant(self, X):
        return X

    def bytes_per_eval(self) -> int:
        return int(sum(X.shape[0] * (self.n_features * X.element_size() + 4) for X in self.Xs))

    def flops_per_eval(self) -> int:
        return int(4 * self.n_rows * self.n_features * self.n_chains)


def synth_logistic_shard(n_rows: int, n_features: int, *, seed: int, device, chunk_rows: int = 1 << 20,
                         beta_scale: float = 0.05):
    """Synthetic logistic-regression shard generated on the device in chunks
    (bf16 ``X ~ N(0,1)``, ``y ~ Bernoulli(sigmoid(X beta* + 0.3))``)."""
    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    beta_true = (torch.randn(n_features, generator=gen, device=device) * beta_scale).float()
    X = torch.empty(n_rows, n_features, dtype=torch.bfloat16, device=device)
    y = torch.empty(n_rows, dtype=torch.float32, device=device)
    for r0 in range(0, n_rows, chunk_rows):
        r1 = min(n_rows, r0 + chunk_rows)
        xb = torch.randn(r1 - r0, n_features, generator=gen, device=device, dtype=torch.float32).to(torch.bfloat16)
        X[r0:r1] = xb
        p = torch.sigmoid(xb.float() @ beta_true + 0.3)
        y[r0:r1] = (torch.rand(r1 - r0, generator=gen, device=device) < p).float()
    return X, y, beta_true
