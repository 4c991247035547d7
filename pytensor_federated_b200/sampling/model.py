"""A minimal probabilistic-model builder on the graph IR (stand-in for ``pm.Model``).

Only what the federated demos need: free scalar/vector variables with Normal priors, arbitrary
potentials (e.g. the ``logp`` output of a ``LogpGradOp``), and compilation of the joint
``logp`` + gradient over one flat vector — the function NUTS and ``find_map`` consume.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .._graph_backend import at, function, grad

LOG_SQRT_2PI = 0.91893853320467274178


class Model:
    def __init__(self) -> None:
        self.free: List[Tuple[str, object, Tuple[int, ...]]] = []   # (name, variable, shape)
        self.terms: List[object] = []
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
        total, *grads = fn(*self.split(theta))
        return float(total), np.concatenate([np.asarray(g, dtype=np.float64).reshape(-1) for g in grads])

    def names(self) -> List[str]:
        out = []
        for name, _, shape in self.free:
            out.extend([name] if not shape else [f"{name}[{i}]" for i in range(shape[0])])
        return out

    def point(self, theta: np.ndarray) -> Dict[str, np.ndarray]:
        return {name: val for (name, _, _), val in zip(self.free, self.split(theta))}
