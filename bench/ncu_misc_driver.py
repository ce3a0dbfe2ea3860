#!/usr/bin/env python
"""Single-GPU workload for `ncu` covering the kernels the other two drivers do not launch: fp8 GEMM +
quantisation, RoPE, embedding fwd/bwd, the fused outer-step kernel (world = 1: all three phases and
both barriers run, the peer loop has one member), delta / master helpers and the peer pull.
  ncu --set full --clock-control none --import-source on -k regex:'fp8|2cta_kernel|rope|embedding|diloco_outer|make_delta|bf16_to_master|peer_pull|amax|colsum' \
      -c 14 -o gpurun_out/prof_misc python bench/ncu_misc_driver.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypha_b200 import ops  # noqa: E402
from hypha_b200.ops import functional as Fn  # noqa: E402
from hypha_b200.ops.native import load_C  # noqa: E402
from hypha_b200.parallel.diloco import AdamWConfig, DiLoCoConfig, DiLoCoOuter, FlatAdamW  # noqa: E402
from hypha_b200.parallel.peer_memory import PeerGroup  # noqa: E402

C = load_C()
s = torch.cuda.current_stream().cuda_stream
bf = dict(device="cuda", dtype=torch.bfloat16)
T, h, inter = 8192, 4096, 14336

# fp8: amax + quantise + 2-CTA kind::f8f6f4 GEMM (gate_up.fwd shape)
x = torch.randn(T, h, **bf)
w = torch.randn(2 * inter, h, **bf) * 0.02
xq, xdq = ops.quantize_fp8(x, ops.E4M3)
wq, wdq = ops.quantize_fp8(w, ops.E4M3)
y = ops.gemm_fp8(xq, xdq, wq, wdq)

# RoPE on the fused qkv activation (32 q + 8 kv heads)
B, S, Hq, Hk, D = 2, 4096, 32, 8, 128
qkv = torch.randn(B * S, (Hq + 2 * Hk) * D, **bf)
cos, sin = Fn.rope_tables(S, D, 500000.0, "cuda")
Fn.rope_qkv(qkv, cos, sin, S, Hq + Hk, D)

# embedding forward / backward at the Llama-3 vocabulary
table = (torch.randn(128256, h, **bf) * 0.02).requires_grad_(True)
ids = torch.randint(0, 128256, (B, S), device="cuda")
e = ops.embedding(ids, table)
e.backward(torch.randn_like(e))

# fused outer step, one rank: 256 Mi parameters (512 MB of delta)
n = 1 << 28
grp = PeerGroup.solo()
pbuf = grp.alloc("p", n, torch.bfloat16)
gbuf = grp.alloc("g", n, torch.bfloat16)
pbuf.local.normal_()
opt = FlatAdamW(pbuf.local, gbuf.local, AdamWConfig())
outer = DiLoCoOuter(opt, DiLoCoConfig(backend="fused", pipeline_chunks=4), grp, pbuf, gbuf)
opt.master.add_(0.01)
outer.step()
assert outer.check() == (0, 0)
C.diloco_make_delta(opt.master.data_ptr(), outer.anchor_buf.local.data_ptr(), gbuf.local.data_ptr(), n, s)
C.bf16_to_master(pbuf.local.data_ptr(), opt.master.data_ptr(), n, s)
grp.pull(0, "p")
torch.cuda.synchronize()
print("ncu misc driver done")
