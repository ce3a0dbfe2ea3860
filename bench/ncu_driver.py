#!/usr/bin/env python
"""Tiny single-GPU workload for `ncu`: a few launches of every hand-written kernel at
Llama-3-8B shapes (T = 8192).  Usage (under gpurun, ONE GPU):
  ncu --set full --clock-control none --import-source on -k regex:<pat> -s <skip> -c <n> \
      -o gpurun_out/prof_<name> python bench/ncu_driver.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypha_b200 import ops  # noqa: E402
from hypha_b200.ops.native import load_C  # noqa: E402

C = load_C()
s = torch.cuda.current_stream().cuda_stream
T, h, inter = 8192, 4096, 14336
bf = dict(device="cuda", dtype=torch.bfloat16)
x = torch.randn(T, h, **bf)
w_gu = torch.randn(2 * inter, h, **bf) * 0.02
for mode in (2, 1):
    C.gemm_set_cta_mode(mode)
    for _ in range(4):
        gu = ops.gemm(x, w_gu)                                   # gate_up.fwd  8192 x 28672 x 4096
dy = torch.randn(T, 2 * inter, **bf)
C.gemm_set_cta_mode(2)
for _ in range(3):
    ops.gemm(dy, w_gu, a_kmajor=True, b_kmajor=False)            # dgrad (MN-major B)
    ops.gemm(dy, x, a_kmajor=False, b_kmajor=False)              # wgrad (both MN-major)
r = torch.randn_like(x)
nw = torch.ones(h, **bf)
y, ro = torch.empty_like(x), torch.empty_like(x)
rstd = torch.empty(T, device="cuda")
nparts = C.rmsnorm_bwd_num_partials(T)
part = torch.empty(nparts, h, device="cuda")
dx = torch.empty_like(x)
act = torch.empty(T, inter, **bf)
dgu = torch.empty_like(gu)
for _ in range(3):
    C.rmsnorm_fwd(x.data_ptr(), r.data_ptr(), ro.data_ptr(), nw.data_ptr(), y.data_ptr(), rstd.data_ptr(), T, h, 1e-5, s)
    C.rmsnorm_bwd(y.data_ptr(), x.data_ptr(), nw.data_ptr(), rstd.data_ptr(), r.data_ptr(), dx.data_ptr(),
                  part.data_ptr(), T, h, s)
    C.swiglu_fwd(gu.data_ptr(), act.data_ptr(), T, inter, s)
    C.swiglu_bwd(act.data_ptr(), gu.data_ptr(), dgu.data_ptr(), T, inter, s)
n = 1 << 28
p = torch.zeros(n, **bf)
g = torch.zeros(n, **bf)
master, theta, mom = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
m, v = torch.zeros(n, **bf), torch.zeros(n, **bf)
for _ in range(3):
    C.adamw_flat(p.data_ptr(), g.data_ptr(), master.data_ptr(), m.data_ptr(), v.data_ptr(), 1, n, 1e-4, 0.9, 0.999,
                 1e-8, 0.01, 0.1, 0.001, 1.0, 7, s)
    C.diloco_outer_local(master.data_ptr(), 0, 0, theta.data_ptr(), mom.data_ptr(), master.data_ptr(), p.data_ptr(),
                         n, 0.7, 0.9, 0, 1.0, s)
logits = torch.randn(2048, 128256, **bf)
tgt = torch.randint(0, 128256, (2048,), device="cuda")
for _ in range(2):
    ops.cross_entropy_inplace(logits, tgt, 1.0 / 2048)
torch.cuda.synchronize()
print("ncu driver done")
