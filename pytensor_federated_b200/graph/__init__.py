"""In-repo symbolic graph IR (PyTensor API subset); see :mod:`.core`."""
from . import core
from .core import (
    Apply,
    Constant,
    DisconnectedType,
    FromFunctionOp,
    FunctionGraph,
    GraphRewriter,
    MergeOptimizer,
    Mode,
    Op,
    ReplaceValidate,
    TensorType,
    Variable,
    apply_depends_on,
    as_tensor,
    function,
    grad,
    optdb,
    scalar,
    vector,
)

__all__ = [
    "core", "Apply", "Constant", "DisconnectedType", "FromFunctionOp", "FunctionGraph", "GraphRewriter",
    "MergeOptimizer", "Mode", "Op", "ReplaceValidate", "TensorType", "Variable", "apply_depends_on",
    "as_tensor", "function", "grad", "optdb", "scalar", "vector",
]
