#!/usr/bin/env python
"""Single-GPU workload for `ncu`: gate_up.fwd (8192 x 28672 x 4096) on the per-tensor fp8 kernel and on the MXFP8
block-scaled kernel, plus the MXFP8 quantiser.  Usage (under gpurun, ONE GPU):
  ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05_2cta|quant_mxfp8" -c 6 \
      -o gpurun_out/prof_fp8 python bench/ncu_fp8_driver.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypha_b200 import ops  # noqa: E402

T, h, inter = 8192, 4096, 14336
x = torch.randn(T, h, device="cuda", dtype=torch.bfloat16)
w = (torch.randn(2 * inter, h, device="cuda") * 0.02).to(torch.bfloat16)
out = torch.empty(T, 2 * inter, device="cuda", dtype=torch.bfloat16)
xq, xdq = ops.quantize_fp8(x, ops.E4M3)
wq, wdq = ops.quantize_fp8(w, ops.E4M3)
for _ in range(2):
    ops.gemm_fp8(xq, xdq, wq, wdq, out=out)
mxq, mxs, mxT, mxTs = ops.quantize_mxfp8(x, rowwise=True, transposed=True)
mwq, mws, _, _ = ops.quantize_mxfp8(w)
for _ in range(2):
    ops.gemm_mxfp8(mxq, mxs, mwq, mws, out=out)
torch.cuda.synchronize()
print("ncu fp8 driver done")
