#!/usr/bin/env python
"""Single-GPU workload for `ncu`: the tcgen05 attention forward at B=2,S=4096 (32 q / 8 kv heads, D=128)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypha_b200 import ops  # noqa: E402

B, S, Hq, Hk, D = 2, 4096, 32, 8, 128
qkv = torch.randn(B * S, (Hq + 2 * Hk) * D, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    out, lse = ops.attention_fwd_tcgen05(qkv, B, S, Hq, Hk, D)
torch.cuda.synchronize()
print("done")
