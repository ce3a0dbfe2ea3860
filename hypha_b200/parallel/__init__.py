"""Parallelism: DiLoCo (inner fused AdamW, outer fused NVLink step) + peer memory."""
from .diloco import AdamWConfig, DiLoCoConfig, DiLoCoOuter, FlatAdamW, make_lr_lambda, shard_bounds
from .peer_memory import BandwidthCounters, PeerGroup, RebindError, SymmetricBuffer, store_exchange

__all__ = ["AdamWConfig", "DiLoCoConfig", "DiLoCoOuter", "FlatAdamW", "make_lr_lambda",
           "shard_bounds", "BandwidthCounters", "PeerGroup", "RebindError", "SymmetricBuffer", "store_exchange"]
