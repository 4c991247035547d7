"""B200-native federated log-probability / gradient engine.

Public surface equals the reference's façade
(``/root/reference/pytensor_federated/__init__.py:1-22``); the PyTensor Ops are
import-guarded so that nodes do not need PyTensor, as in the reference.
"""
try:
    from .op_async import AsyncOp
    from .wrapper_ops import (
        ArraysToArraysOp,
        AsyncArraysToArraysOp,
        AsyncLogpGradOp,
        AsyncLogpOp,
        LogpGradOp,
        LogpOp,
    )
except ModuleNotFoundError:
    pass
from .common import (
    LogpGradServiceClient,
    LogpServiceClient,
    wrap_logp_func,
    wrap_logp_grad_func,
)
from .service import ArraysToArraysService, ArraysToArraysServiceClient
from .signatures import ComputeFunc, LogpFunc, LogpGradFunc

__version__ = "0.1.0"
