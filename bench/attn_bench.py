#!/usr/bin/env python
"""Attention forward + backward micro-benchmark at the Llama-3-8B shape (B=2, S=4096, 32 q / 8 kv heads,
D=128, causal): hand-written tcgen05 kernel vs cuDNN (torch SDPA) vs flash-attn 2.
FLOPs counted as 4*B*H*S^2*D/2 (causal).  -> gpurun_out/attn_bench.json"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypha_b200 import ops  # noqa: E402
from hypha_b200.ops.functional import _attention_sdpa, _FlashAttnQKV  # noqa: E402


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    res = []
    for B, S in ((2, 4096), (1, 8192), (4, 2048)):
        Hq, Hk, D = 32, 8, 128
        qkv = torch.randn(B * S, (Hq + 2 * Hk) * D, device="cuda", dtype=torch.bfloat16)
        flops = 4.0 * B * Hq * S * S * D / 2
        row = {"B": B, "S": S}
        with torch.no_grad():
            for name, fn in (("tcgen05", lambda: ops.attention_fwd_tcgen05(qkv, B, S, Hq, Hk, D)),
                             ("tcgen05_v3", lambda: ops.attention_fwd_tcgen05(qkv, B, S, Hq, Hk, D, version=3)),
                             ("tcgen05_v2", lambda: ops.attention_fwd_tcgen05(qkv, B, S, Hq, Hk, D, version=2)),
                             ("tcgen05_v1", lambda: ops.attention_fwd_tcgen05(qkv, B, S, Hq, Hk, D, version=1)),
                             ("cudnn", lambda: _attention_sdpa(qkv, B, S, Hq, Hk, D)),
                             ("flash2", lambda: _FlashAttnQKV.apply(qkv, B, S, Hq, Hk, D))):
                try:
                    ms = timeit(fn)
                    row[name] = {"ms": ms, "tflops": flops / ms / 1e9}
                except Exception as exc:
                    row[name] = {"error": repr(exc)[:200]}
            for poly in (0, 1, 2):
                ops.set_attention_fwd_poly(poly)
                ms = timeit(lambda: ops.attention_fwd_tcgen05(qkv, B, S, Hq, Hk, D))
                a, _ = ops.attention_fwd_tcgen05(qkv, B, S, Hq, Hk, D)
                bb = _attention_sdpa(qkv, B, S, Hq, Hk, D)
                row[f"tcgen05_poly{poly}"] = {"ms": ms, "tflops": flops / ms / 1e9,
                                              "rel_err_vs_cudnn": ((a.float() - bb.float()).norm() / bb.float().norm()).item()}
            ops.set_attention_fwd_poly(1)
            a, lse = ops.attention_fwd_tcgen05(qkv, B, S, Hq, Hk, D)
            b = _attention_sdpa(qkv, B, S, Hq, Hk, D)
            row["rel_err_vs_cudnn"] = ((a.float() - b.float()).norm() / b.float().norm()).item()
            # ---- backward (2.5x the forward FLOPs): hand-written tcgen05 vs library fwd+bwd minus fwd
            dout = torch.randn_like(a)
            for variant in ("fused", "split"):
                try:
                    ms = timeit(lambda: ops.attention_bwd_tcgen05(qkv, a, dout, lse, B, S, Hq, Hk, D, variant=variant))
                    row["bwd_tcgen05_" + variant] = {"ms": ms, "tflops": 2.5 * flops / ms / 1e9}
                except Exception as exc:
                    row["bwd_tcgen05_" + variant] = {"error": repr(exc)[:200]}
            try:
                d1 = ops.attention_bwd_tcgen05(qkv, a, dout, lse, B, S, Hq, Hk, D, variant="fused").float()
                d2 = ops.attention_bwd_tcgen05(qkv, a, dout, lse, B, S, Hq, Hk, D, variant="split").float()
                row["bwd_fused_vs_split_rel"] = ((d1 - d2).norm() / d2.norm()).item()
            except Exception as exc:
                row["bwd_fused_vs_split_rel"] = repr(exc)[:200]
        for name in ("cudnn", "flash2", "tcgen05"):
            # library backward = (forward + backward through the autograd Function) - forward
            try:
                ops.set_attention_backend(name)
                q = qkv.clone().requires_grad_(True)
                g = torch.randn(B * S, Hq * D, device="cuda", dtype=torch.bfloat16)

                def fb():
                    q.grad = None
                    ops.attention_qkv(q, B, S, Hq, Hk, D).backward(g)

                ms = timeit(fb)
                row["fwd+bwd_" + name] = {"ms": ms, "tflops": 3.5 * flops / ms / 1e9}
                if name != "tcgen05" and "ms" in row.get(name, {}):
                    bms = ms - row[name]["ms"]
                    row["bwd_" + name] = {"ms": bms, "tflops": 2.5 * flops / bms / 1e9}
            except Exception as exc:
                row["fwd+bwd_" + name] = {"error": repr(exc)[:200]}
        ops.set_attention_backend("cudnn")
        res.append(row)
        print(json.dumps(row), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/attn_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
