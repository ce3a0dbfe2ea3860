"""Small shared helpers."""
from __future__ import annotations

import os
import random
from typing import Optional


def seed_everything(seed: int) -> None:
    import numpy as np
    import torch

    random.seed(seed)
    np.random.seed(seed % (2 ** 32))
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def human_bytes(n: float) -> str:
    for unit in ("B", "KiB", "MiB", "GiB", "TiB"):
        if abs(n) < 1024.0:
            return f"{n:.1f} {unit}"
        n /= 1024.0
    return f"{n:.1f} PiB"


def env_flag(name: str, default: bool = False) -> bool:
    v = os.environ.get(name)
    return default if v is None else v.strip().lower() in ("1", "true", "yes", "on")


def memory_plan_gb(num_params: int, world: int, moments_bytes: int = 2) -> dict:
    """Bytes (decimal GB) of persistent training state per rank for the DiLoCo trainer
    (docs/training.md): the planning rule the reference states as "GB ~ params x 24 x 9e-10"
    (docs/scheduler.md:149) made explicit for the bf16 + sharded-outer-state layout."""
    p = num_params / 1e9
    plan = {"params_bf16": 2 * p, "grads_bf16": 2 * p, "master_fp32": 4 * p, "adam_moments": 2 * moments_bytes * p,
            "outer_theta_momentum_shard": 8 * p / world, "anchor_bf16": (2 * p if world > 1 else 0.0)}
    plan["total"] = sum(plan.values())
    return plan
