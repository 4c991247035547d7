"""Model families with fused sm_100a kernels (and eager PyTorch oracles)."""
from .base import ShardModel
from .glm import GlmShards, synth_logistic_shard
from .linreg import LinregShards, make_demo_data
from .ode import OdeShards, synth_lv_shard

__all__ = [
    "ShardModel",
    "LinregShards",
    "make_demo_data",
    "GlmShards",
    "synth_logistic_shard",
    "OdeShards",
    "synth_lv_shard",
]
