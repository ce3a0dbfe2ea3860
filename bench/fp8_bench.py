#!/usr/bin/env python
"""fp8 GEMM microbenchmark: Llama-3-8B layer GEMMs at T tokens in the three Linear-layer roles, all K-major:
bf16 tcgen05 2-CTA vs per-tensor fp8 (kind::f8f6f4) vs MXFP8 (kind::mxf8f6f4.block_scale), plus the one-pass
MXFP8 quantiser.  CUDA events, warm-up, operands rotated through > L2.  Writes ``gpurun_out/fp8_bench.json``."""
from __future__ import annotations

import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypha_b200 import ops  # noqa: E402
from hypha_b200.ops.native import load_C  # noqa: E402

from kernel_bench import peaks, time_fn  # noqa: E402


def case(name, M, N, K, out_rows):
    dev = "cuda"
    nvar = max(1, min(6, int(300e6 // max(1, (M * K + N * K) + M * N * 2)) + 1))
    A = [torch.randn(M, K, device=dev, dtype=torch.bfloat16) for _ in range(nvar)]
    B = [(torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16) for _ in range(nvar)]
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    row = {"name": name, "M": M, "N": N, "K": K}
    load_C().gemm_set_cta_mode(2)
    ms = time_fn(lambda i: ops.gemm(A[i], B[i], out=out), variants=nvar)
    row["bf16"] = {"ms": ms, "tflops": 2.0 * M * N * K / ms / 1e9}
    qa = [ops.quantize_fp8(a, ops.E4M3) for a in A]
    qb = [ops.quantize_fp8(b, ops.E4M3) for b in B]
    ms = time_fn(lambda i: ops.gemm_fp8(qa[i][0], qa[i][1], qb[i][0], qb[i][1], out=out), variants=nvar)
    row["fp8_tensor"] = {"ms": ms, "tflops": 2.0 * M * N * K / ms / 1e9}
    ma = [ops.quantize_mxfp8(a)[:2] for a in A]
    mb = [ops.quantize_mxfp8(b)[:2] for b in B]
    ms = time_fn(lambda i: ops.gemm_mxfp8(ma[i][0], ma[i][1], mb[i][0], mb[i][1], out=out), variants=nvar)
    row["mxfp8"] = {"ms": ms, "tflops": 2.0 * M * N * K / ms / 1e9}
    ref = ops.mxfp8_dequant_ref(*ma[0])[:512] @ ops.mxfp8_dequant_ref(*mb[0]).t()
    got = ops.gemm_mxfp8(ma[0][0], ma[0][1], mb[0][0], mb[0][1])[:512].float()
    row["mxfp8_rel_err_vs_dequantised_fp32"] = ((got - ref).norm() / ref.norm()).item()
    row["mxfp8_vs_bf16"] = row["bf16"]["ms"] / row["mxfp8"]["ms"]
    out_rows.append(row)
    print(json.dumps(row), flush=True)


def main():
    hbm, tf, src = peaks()
    res = {"peaks": {"hbm_gbs": hbm, "bf16_tflops": tf, "source": src}, "gemm": [], "quantize": []}
    T = int(os.environ.get("KB_TOKENS", "8192"))
    h, inter, qkv = 4096, 14336, 6144
    for nm, n_out, k_in in [("qkv", qkv, h), ("wo", h, h), ("gate_up", 2 * inter, h), ("down", h, inter)]:
        case(f"{nm}.fwd", T, n_out, k_in, res["gemm"])
        case(f"{nm}.dgrad", T, k_in, n_out, res["gemm"])
        case(f"{nm}.wgrad", n_out, k_in, T, res["gemm"])
    for R, Cc in [(T, h), (T, 2 * inter), (2 * inter, h)]:
        x = torch.randn(R, Cc, device="cuda", dtype=torch.bfloat16)
        for row_, col_, tag, out_b in ((True, False, "row", 1), (False, True, "transposed", 1), (True, True, "both", 2)):
            ms = time_fn(lambda i: ops.quantize_mxfp8(x, rowwise=row_, transposed=col_))
            b = R * Cc * (2 + out_b)
            res["quantize"].append({"shape": [R, Cc], "outputs": tag, "ms": ms, "gbs": b / ms / 1e6,
                                    "frac_of_hbm": b / ms / 1e6 / hbm})
            print(json.dumps(res["quantize"][-1]), flush=True)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    json.dump(res, open(os.path.join(out, "fp8_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
