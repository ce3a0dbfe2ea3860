#!/usr/bin/env python
"""Kernel-time breakdown of one Llama-3-8B inner step (torch.profiler / CUPTI, not ncu).

Shares of device time per kernel family -> gpurun_out/profile_<tag>.json.  Absolute times under
CUPTI are slightly inflated; compare SHARES."""
import argparse
import json
import os
import sys
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypha_b200 import ops  # noqa: E402
from hypha_b200.models.llama import get_config  # noqa: E402
from hypha_b200.parallel.diloco import AdamWConfig, DiLoCoConfig  # noqa: E402
from hypha_b200.trainer import DiLoCoTrainer  # noqa: E402


def family(name: str) -> str:
    n = name.lower()
    for key, fam in (("attn_fwd", "attention.fwd"), ("attn_bwd", "attention.bwd"), ("gemm_tcgen05_2cta", "gemm.tcgen05"),
                     ("gemm_bf16_tcgen05", "gemm.tcgen05"), ("nvjet", "gemm.cublas"), ("cutlass", "gemm.cublas"),
                     ("gemm", "gemm.cublas"), ("flash_fwd", "attention.fwd"), ("flash_bwd", "attention.bwd"),
                     ("flash", "attention.other"), ("rmsnorm_fwd", "rmsnorm.fwd"), ("rmsnorm_bwd", "rmsnorm.bwd"),
                     ("colsum", "rmsnorm.bwd"), ("rope", "rope"), ("swiglu_fwd", "swiglu.fwd"),
                     ("swiglu_bwd", "swiglu.bwd"), ("cross_entropy", "cross_entropy"), ("adamw", "adamw"),
                     ("embedding", "embedding"), ("outer", "outer_step"), ("memcpy", "memcpy"), ("memset", "memset")):
        if key in n:
            return fam
    return "other:" + name[:60]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="tcgen05")
    ap.add_argument("--model", default="llama3-8b")
    ap.add_argument("--seq-len", type=int, default=4096)
    ap.add_argument("--micro-batch", type=int, default=2)
    ap.add_argument("--recompute-layers", type=int, default=0)
    ap.add_argument("--tag", default=None)
    a = ap.parse_args()
    ops.set_gemm_backend(a.backend)
    cfg = get_config(a.model, recompute_layers=a.recompute_layers)
    tr = DiLoCoTrainer.create(cfg, adamw=AdamWConfig(), diloco=DiLoCoConfig(inner_steps=1000), device=0)
    ids = torch.randint(0, cfg.vocab_size, (a.micro_batch, a.seq_len), device="cuda")
    for _ in range(3):
        tr.step(ids)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        tr.step(ids)
        torch.cuda.synchronize()
    fam = defaultdict(lambda: [0.0, 0])
    kernels = []
    total = 0.0
    for ev in prof.key_averages():
        t = getattr(ev, "device_time_total", None)
        if t is None:
            t = getattr(ev, "cuda_time_total", 0.0)
        if t <= 0:
            continue
        f = family(ev.key)
        kernels.append({"kernel": ev.key[:100], "ms": t / 1e3, "launches": ev.count})
        fam[f][0] += t / 1e3
        fam[f][1] += ev.count
        total += t / 1e3
    rows = sorted(((k, v[0], v[1]) for k, v in fam.items()), key=lambda r: -r[1])
    out = {"backend": a.backend, "config": vars(a), "total_kernel_ms": total,
           "families": [{"family": k, "ms": ms, "share": ms / total, "launches": n} for k, ms, n in rows],
           "kernels": sorted(kernels, key=lambda r: -r["ms"])[:40]}
    tag = a.tag or a.backend
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/profile_{tag}.json", "w"), indent=1)
    print(f"total kernel time {total:.1f} ms")
    for k, ms, n in rows[:24]:
        print(f"{ms:9.2f} ms {100 * ms / total:5.1f}%  x{n:<5d} {k}")


if __name__ == "__main__":
    main()
