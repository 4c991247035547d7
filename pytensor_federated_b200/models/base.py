"""Contract between a model family and the federation engine.

A *shard model* is the B200-native counterpart of the black-box ``compute_func`` that a
reference node serves (``/root/reference/pytensor_federated/service.py:78-86``): it owns the
node's private data (resident in HBM) and knows

* how the client's input arrays are packed into the theta mailbox (32-bit words),
* how the reduced ``[LL, dLL/dtheta ...]`` vector of doubles is unpacked into the flat
  ``(logp, *gradients)`` tuple of the ``wrap_logp_grad_func`` convention
  (``/root/reference/pytensor_federated/common.py:26-49``),
* how to attach itself to a native engine (which kernel, which launch shape), and
* an eager PyTorch implementation of the same maths (numerics oracle for the kernels and the
  compute step of the NCCL/gloo baseline path).
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


class ShardModel:
    """Base class; subclasses set ``n_theta_words`` and ``n_vals``."""

    #: number of 32-bit words in the theta mailbox
    n_theta_words: int = 0
    #: number of float64 values in a node partial / the reduced result
    n_vals: int = 0
    #: number of input arrays of the ArraysToArrays signature
    n_inputs: int = 0

    def pack_theta(self, inputs: Sequence[np.ndarray], out: np.ndarray) -> None:
        """Writes the inputs into ``out`` (``uint32[n_theta_words]`` view of pinned memory)."""
        raise NotImplementedError

    def unpack_result(self, vals: np.ndarray) -> List[np.ndarray]:
        """``float64[n_vals]`` → ``[logp, grad_0, grad_1, ...]`` (fresh arrays)."""
        raise NotImplementedError

    def attach(self, lib, handle) -> None:
        """Registers data pointers and kernel choice with the native engine."""
        raise NotImplementedError

    def reference_partial(self, inputs: Sequence[np.ndarray]) -> np.ndarray:
        """This node's partial ``float64[n_vals]`` computed with stock PyTorch ops."""
        raise NotImplementedError

    # -- conveniences --------------------------------------------------------------------
    def reference(self, inputs: Sequence[np.ndarray]) -> List[np.ndarray]:
        return self.unpack_result(self.reference_partial(inputs))

    def bytes_per_eval(self) -> int:
        """Algorithmic HBM bytes one evaluation must move on this node (roofline input)."""
        return 0

    def flops_per_eval(self) -> int:
        return 0
