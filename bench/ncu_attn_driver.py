#!/usr/bin/env python
"""Single-GPU workload for `ncu`: the hand-written attention kernels at the Llama-3-8B shape
(B=2, S=4096, 32 q / 8 kv heads, D=128, causal), two rounds of forward + backward.
  ncu --set full --clock-control none --import-source on -k regex:attn_ -s 4 -c 4 \
      -o gpurun_out/prof_attn python bench/ncu_attn_driver.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypha_b200 import ops  # noqa: E402

B, S, Hq, Hk, D = 2, 4096, 32, 8, 128
qkv = torch.randn(B * S, (Hq + 2 * Hk) * D, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    out, lse = ops.attention_fwd_tcgen05(qkv, B, S, Hq, Hk, D)
    dout = torch.randn_like(out)
    dqkv = ops.attention_bwd_tcgen05(qkv, out, dout, lse, B, S, Hq, Hk, D)
torch.cuda.synchronize()
print("ncu attention driver done")
