"""Multi-GPU execution: the fused NVLink engine and the collective (NCCL/gloo) baseline."""
from .engine import EngineClosedError, FederatedEngine, FederationError, FederationTimeout

__all__ = ["FederatedEngine", "FederationError", "FederationTimeout", "EngineClosedError"]
