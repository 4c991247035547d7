"""A minimal probabilistic-model builder on the graph IR (stand-in for ``pm.Model``).

Only what the federated demos need: free scalar/vector variables with Normal priors, arbitrary
potentials (e.g. the ``logp`` output of a ``LogpGradOp``), and compilation of the joint
``logp`` + gradient over one flat vector — the function NUTS and ``find_map`` consume.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np

from .._graph_backend import at, function, grad

LOG_SQRT_2PI = 0.91893853320467274178


class Model:
    """Minimal model builder over the graph backend: free variables with priors (``Flat``, ``Normal``,
    ``HalfNormal``, ``Exponential``, ``Uniform`` — constrained ones are sampled on an unconstrained scale),
    ``Potential`` terms (e.g. the outputs of ``LogpGradOp`` nodes), a compiled ``logp_dlogp(theta)`` and the
    ``find_map`` / ``sample`` conveniences — what ``pm.Model`` is to the reference's ``demo_model.py:28-44``."""

    def __init__(self) -> None:
        self.free: List[Tuple[str, object, Tuple[int, ...]]] = []   # (name, variable, shape)
        self.terms: List[object] = []
        self.transforms: Dict[str, Tuple[str, object]] = {}   # constrained name -> (free name, back-transform)
        self._compiled = None

    # -- building ----------------------------------------------------------------------------
    def _new(self, name: str, size: Optional[int]):
        var = at.scalar(name) if size is None else at.vector(name)
        self.free.append((name, var, () if size is None else (int(size),)))
        self._compiled = None
        return var

    def Flat(self, name: str, size: Optional[int] = None):
        return self._new(name, size)

    def Normal(self, name: str, mu=0.0, sigma=1.0, size: Optional[int] = None):
        var = self._new(name, size)
        z = (var - mu) / sigma
        n = 1 if size is None else size
        term = (-0.5 * z * z).sum() - n * (np.log(sigma) + LOG_SQRT_2PI) if not hasattr(sigma, "type") else (
            (-0.5 * z * z - at.log(sigma)).sum() - n * LOG_SQRT_2PI
        )
        self.terms.append(term)
        return var

    def HalfNormal(self, name: str, sigma=1.0):
        """Positive scalar with a half-normal prior, sampled on the log scale (Jacobian included).
        Returns the constrained variable; the free parameter is ``<name>_log__``."""
        log_var = self._new(f"{name}_log__", None)
        var = at.exp(log_var)
        z = var / sigma
        # log N+(x | sigma) + log |dx / dlog x| = log 2 - log sigma - log sqrt(2 pi) - z^2 / 2 + log x
        self.terms.append(-0.5 * z * z + log_var + (np.log(2.0) - np.log(sigma) - LOG_SQRT_2PI))
        self.transforms[name] = (f"{name}_log__", np.exp)
        return var

    def Exponential(self, name: str, lam=1.0):
        """Positive scalar with density ``lam * exp(-lam x)``, sampled on the log scale."""
        log_var = self._new(f"{name}_log__", None)
        var = at.exp(log_var)
        self.terms.append(np.log(lam) - lam * var + log_var)           # + log |dx / dlog x|
        self.transforms[name] = (f"{name}_log__", np.exp)
        return var

    def Uniform(self, name: str, lower=0.0, upper=1.0):
        """Scalar on ``(lower, upper)`` with a flat prior, sampled on the logit scale."""
        if not upper > lower:
            raise ValueError("upper must exceed lower")
        logit = self._new(f"{name}_interval__", None)
        unit = at.sigmoid(logit)
        var = lower + (upper - lower) * unit
        # density 1 / (upper - lower) times |dx / dlogit| = (upper - lower) * s * (1 - s)
        # log s + log(1 - s) in the overflow-free form
        self.terms.append(-(at.softplus(-logit) + at.softplus(logit)))
        self.transforms[name] = (f"{name}_interval__", lambda z, lo=lower, hi=upper: lo + (hi - lo) / (1.0 + np.exp(-z)))
        return var

    def Potential(self, name: str, var) -> None:
        self.terms.append(at.as_tensor(var).sum() if getattr(var, "ndim", 0) else var)
        self._compiled = None

    # -- compilation -------------------------------------------------------------------------
    @property
    def dim(self) -> int:
        return int(sum(int(np.prod(s)) if s else 1 for _, _, s in self.free))

    def _total(self):
        total = self.terms[0]
        for t in self.terms[1:]:
            total = total + t
        return total

    def compile(self, mode=None):
        variables = [v for _, v, _ in self.free]
        total = self._total()
        grads = grad(total, variables)
        fn = function(variables, [total, *grads], mode=mode)
        self._compiled = fn
        return fn

    def compile_logp(self, mode=None):
        """The joint log-probability WITHOUT gradients — all a model with gradient-free potentials
        (``LogpOp`` / ``AsyncLogpOp``, which define no ``grad``) can offer; Metropolis consumes it."""
        fn = function([v for _, v, _ in self.free], [self._total()], mode=mode)
        self._compiled_logp = fn
        return fn

    def logp(self, theta: np.ndarray) -> float:
        fn = getattr(self, "_compiled_logp", None) or self.compile_logp()
        (total,) = fn(*self.split(np.atleast_1d(theta)))
        return float(total)

    def split(self, theta: np.ndarray) -> List[np.ndarray]:
        out, pos = [], 0
        for _, _, shape in self.free:
            n = int(np.prod(shape)) if shape else 1
            chunk = np.asarray(theta[pos : pos + n], dtype=np.float64)
            out.append(chunk.reshape(shape) if shape else chunk.reshape(()))
            pos += n
        return out

    def logp_dlogp(self, theta: np.ndarray):
        fn = self._compiled or self.compile()
        # (start, stop, shape) of every free variable inside the flat vector; built once per compilation — this
        # function runs once per leapfrog step
        plan = self.__dict__.get("_flat_plan")
        if plan is None or plan[0] is not fn:
            spans, pos = [], 0
            for _, _, shape in self.free:
                n = int(np.prod(shape)) if shape else 1
                spans.append((pos, pos + n, tuple(shape) if shape else ()))
                pos += n
            plan = self._flat_plan = (fn, spans, pos)
        _, spans, dim = plan
        theta = np.asarray(theta, dtype=np.float64).reshape(-1)
        if theta.size != dim:
            raise ValueError(f"expected a parameter vector of length {dim}, got {theta.size}")
        total, *grads = fn(*[theta[a:b].reshape(shape) for a, b, shape in spans])
        out = np.empty(dim)
        for (a, b, _), g in zip(spans, grads):
            out[a:b] = np.reshape(g, -1)
        return float(total), out

    def names(self) -> List[str]:
        out = []
        for name, _, shape in self.free:
            out.extend([name] if not shape else [f"{name}[{i}]" for i in range(shape[0])])
        return out

    def point(self, theta: np.ndarray) -> Dict[str, np.ndarray]:
        out = {name: val for (name, _, _), val in zip(self.free, self.split(theta))}
        for name, (free_name, back) in self.transforms.items():
            out[name] = back(out[free_name])
        return out

    # -- inference conveniences (what pm.find_MAP / pm.sample are to a pm.Model) ----------------
    def find_map(self, start: Optional[np.ndarray] = None, **kwargs):
        from .mcmc import find_map

        theta, info = find_map(self.logp_dlogp, np.zeros(self.dim) if start is None else start, **kwargs)
        return self.point(theta), info

    def sample_metropolis(self, draws: int = 1000, tune: int = 1000, *, chains: int = 1,
                          start: Optional[np.ndarray] = None, seed: int = 0, **kwargs):
        """Random-walk Metropolis on :meth:`logp` — the sampler for models whose potentials have no gradient
        (what ``pm.sample(step=pm.Metropolis())`` is to the reference's ``LogpOp`` test).  Returns like
        :meth:`sample`."""
        from .mcmc import metropolis_sample

        start = np.zeros(self.dim) if start is None else np.asarray(start, dtype=np.float64)
        results = [metropolis_sample(self.logp, start, draws=draws, tune=tune, seed=seed + c, **kwargs) for c in range(chains)]
        return self._columns(results, chains)

    def _columns(self, results, chains: int):
        samples = np.stack([r.samples for r in results], axis=1)          # [draws, chains, dim]
        columns = {}
        pos = 0
        for name, _, shape in self.free:
            n = int(np.prod(shape)) if shape else 1
            block = samples[:, :, pos : pos + n]
            columns[name] = block[:, :, 0] if not shape else block
            pos += n
        for name, (free_name, back) in self.transforms.items():
            columns[name] = back(columns[free_name])
        if chains == 1:
            return results[0], {k: v[:, 0] for k, v in columns.items()}
        return results, columns

    def sample(self, draws: int = 200, tune: int = 500, *, chains: int = 1, start: Optional[np.ndarray] = None,
               seed: int = 0, **kwargs):
        """NUTS from ``start`` (default: the MAP).  ``chains == 1`` returns ``(SamplerResult, {name: draws})``;
        ``chains > 1`` runs them one after the other with seeds ``seed, seed + 1, ...`` and returns
        ``([SamplerResult, ...], {name: draws[draws, chains(, k)]})`` — feed the dict to
        :func:`~pytensor_federated_b200.sampling.summarize` for ESS / R-hat."""
        from .mcmc import find_map, nuts_sample

        if start is None:
            start, _ = find_map(self.logp_dlogp, np.zeros(self.dim))
        results = [nuts_sample(self.logp_dlogp, start, draws=draws, tune=tune, seed=seed + c, **kwargs)
                   for c in range(chains)]
        return self._columns(results, chains)
