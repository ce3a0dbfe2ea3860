#!/usr/bin/env python
"""Per-kernel microbenchmarks (CUDA events, warm-up, L2-cold rotation of operands).

Writes ``gpurun_out/kernel_bench.json``.  GEMM shapes are the Llama-3-8B layer GEMMs at
T = 8192 tokens in the three layouts a Linear layer needs (fwd / dgrad / wgrad), hand-written
tcgen05 kernel vs cuBLAS (torch.mm).  Fractions are of the measured peaks in MEASURED_PEAKS.json.
"""
from __future__ import annotations

import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypha_b200 import ops  # noqa: E402
from hypha_b200.ops.native import load_C  # noqa: E402


def peaks():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    try:
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], "measured"
    except Exception:
        return 6650.0, 1590.0, "fallback"


def time_fn(fn, iters=20, warm=5, variants=1):
    for i in range(warm):
        fn(i % variants)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % variants)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def gemm_case(name, M, N, K, a_k, b_k, results, hbm, tf):
    dev = "cuda"
    nvar = max(1, min(8, int(300e6 // max(1, (M * K + N * K + M * N) * 2)) + 1))  # rotate > L2
    As = [torch.randn((M, K) if a_k else (K, M), device=dev, dtype=torch.bfloat16) for _ in range(nvar)]
    Bs = [torch.randn((N, K) if b_k else (K, N), device=dev, dtype=torch.bfloat16) for _ in range(nvar)]
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    row = {"name": name, "M": M, "N": N, "K": K, "a_kmajor": a_k, "b_kmajor": b_k}
    C = load_C()
    for backend in ("tcgen05", "tcgen05_2cta", "cublas"):
        ops.set_gemm_backend("cublas" if backend == "cublas" else "tcgen05")
        C.gemm_set_cta_mode(2 if backend == "tcgen05_2cta" else 1)
        try:
            ms = time_fn(lambda i: ops.gemm(As[i], Bs[i], a_kmajor=a_k, b_kmajor=b_k, out=out),
                         variants=nvar)
            tflops = 2.0 * M * N * K / ms / 1e9
            row[backend] = {"ms": ms, "tflops": tflops, "frac_of_peak": tflops / tf}
        except Exception as exc:  # keep going: one broken layout must not hide the others
            row[backend] = {"error": repr(exc)}
    # correctness spot check
    try:
        ops.set_gemm_backend("cublas")
        ref = ops.gemm(As[0], Bs[0], a_kmajor=a_k, b_kmajor=b_k).float()
        ops.set_gemm_backend("tcgen05")
        for mode, key in ((1, "rel_err_vs_cublas"), (2, "rel_err_2cta_vs_cublas")):
            C.gemm_set_cta_mode(mode)
            got = ops.gemm(As[0], Bs[0], a_kmajor=a_k, b_kmajor=b_k).float()
            row[key] = ((got - ref).norm() / ref.norm()).item()
        C.gemm_set_cta_mode(1)
    except Exception as exc:
        row["rel_err_vs_cublas"] = repr(exc)
    results.append(row)
    print(json.dumps(row), flush=True)


def main():
    hbm, tf, src = peaks()
    C = load_C()
    res = {"peaks": {"hbm_gbs": hbm, "bf16_tflops": tf, "source": src}, "gemm": [], "elementwise": []}
    T = int(os.environ.get("KB_TOKENS", "8192"))
    h, inter, qkv, V = 4096, 14336, 6144, 128256
    shapes = [("qkv", qkv, h), ("wo", h, h), ("gate_up", 2 * inter, h), ("down", h, inter)]
    for nm, n_out, k_in in ([] if os.environ.get("KB_SKIP_GEMM") == "1" else shapes):
        gemm_case(f"{nm}.fwd", T, n_out, k_in, True, True, res["gemm"], hbm, tf)
        gemm_case(f"{nm}.dgrad", T, k_in, n_out, True, False, res["gemm"], hbm, tf)
        gemm_case(f"{nm}.wgrad", n_out, k_in, T, False, False, res["gemm"], hbm, tf)
    if os.environ.get("KB_SKIP_GEMM") != "1":
        gemm_case("lm_head.fwd(chunk4096)", 4096, V, h, True, True, res["gemm"], hbm, tf)
        gemm_case("square8192", 8192, 8192, 8192, True, True, res["gemm"], hbm, tf)
    # gate_up forward with and without the fused SwiGLU epilogue (vs GEMM + stand-alone SwiGLU kernel)
    xa = torch.randn(T, h, device="cuda", dtype=torch.bfloat16)
    wgu = (0.02 * torch.randn(2 * inter, h, device="cuda")).to(torch.bfloat16)
    gu_o = torch.empty(T, 2 * inter, device="cuda", dtype=torch.bfloat16)
    act_o = torch.empty(T, inter, device="cuda", dtype=torch.bfloat16)
    ms_plain = time_fn(lambda i: ops.gemm(xa, wgu, out=gu_o))
    ms_sep = time_fn(lambda i: (ops.gemm(xa, wgu, out=gu_o), C.swiglu_fwd(gu_o.data_ptr(), act_o.data_ptr(), T, inter,
                                                                          torch.cuda.current_stream().cuda_stream)))
    ms_fused = time_fn(lambda i: ops.gemm(xa, wgu, out=gu_o, swiglu_out=act_o))
    res["swiglu_epilogue"] = {"gemm_ms": ms_plain, "gemm_plus_swiglu_kernel_ms": ms_sep, "fused_ms": ms_fused}
    print(json.dumps(res["swiglu_epilogue"]), flush=True)
    # down-projection dgrad with and without the fused SwiGLU-backward epilogue
    dyd = torch.randn(T, h, device="cuda", dtype=torch.bfloat16)
    wd = (0.02 * torch.randn(h, inter, device="cuda")).to(torch.bfloat16)
    dgu_o = torch.empty_like(gu_o)
    sb = torch.cuda.current_stream().cuda_stream
    ms_plain = time_fn(lambda i: ops.gemm(dyd, wd, a_kmajor=True, b_kmajor=False, out=act_o))
    ms_sep = time_fn(lambda i: (ops.gemm(dyd, wd, a_kmajor=True, b_kmajor=False, out=act_o),
                                C.swiglu_bwd(act_o.data_ptr(), gu_o.data_ptr(), dgu_o.data_ptr(), T, inter, sb)))
    ms_fused = time_fn(lambda i: ops.gemm(dyd, wd, a_kmajor=True, b_kmajor=False, out=act_o, dswiglu=(gu_o, dgu_o)))
    res["dswiglu_epilogue"] = {"dgrad_ms": ms_plain, "dgrad_plus_swiglu_bwd_kernel_ms": ms_sep, "fused_ms": ms_fused}
    print(json.dumps(res["dswiglu_epilogue"]), flush=True)
    ops.set_gemm_backend("tcgen05")

    # ---- memory-bound kernels ---------------------------------------------------------------
    s = torch.cuda.current_stream().cuda_stream
    x = torch.randn(T, h, device="cuda", dtype=torch.bfloat16)
    r = torch.randn_like(x)
    w = torch.ones(h, device="cuda", dtype=torch.bfloat16)
    y, ro = torch.empty_like(x), torch.empty_like(x)
    rstd = torch.empty(T, device="cuda")
    ms = time_fn(lambda i: C.rmsnorm_fwd(x.data_ptr(), r.data_ptr(), ro.data_ptr(), w.data_ptr(), y.data_ptr(),
                                         rstd.data_ptr(), T, h, 1e-5, s))
    bytes_ = 4 * T * h * 2
    res["elementwise"].append({"name": "add_rmsnorm_fwd", "ms": ms, "gbs": bytes_ / ms / 1e6,
                               "frac_of_hbm": bytes_ / ms / 1e6 / hbm})
    nparts = C.rmsnorm_bwd_num_partials(T)
    part = torch.empty(nparts, h, device="cuda")
    dx = torch.empty_like(x)
    ms = time_fn(lambda i: C.rmsnorm_bwd(y.data_ptr(), x.data_ptr(), w.data_ptr(), rstd.data_ptr(), r.data_ptr(),
                                         dx.data_ptr(), part.data_ptr(), T, h, s))
    res["elementwise"].append({"name": "rmsnorm_bwd", "ms": ms, "gbs": bytes_ / ms / 1e6,
                               "frac_of_hbm": bytes_ / ms / 1e6 / hbm})
    wg = torch.zeros(h, device="cuda", dtype=torch.bfloat16)
    ms = time_fn(lambda i: C.colsum_partials(part.data_ptr(), nparts, h, wg.data_ptr(), 0, s))
    res["elementwise"].append({"name": f"colsum_partials[{nparts}x{h}]", "ms": ms})
    gu = torch.randn(T, 2 * inter, device="cuda", dtype=torch.bfloat16)
    act = torch.empty(T, inter, device="cuda", dtype=torch.bfloat16)
    ms = time_fn(lambda i: C.swiglu_fwd(gu.data_ptr(), act.data_ptr(), T, inter, s))
    b = 3 * T * inter * 2
    res["elementwise"].append({"name": "swiglu_fwd", "ms": ms, "gbs": b / ms / 1e6, "frac_of_hbm": b / ms / 1e6 / hbm})
    dgu = torch.empty_like(gu)
    ms = time_fn(lambda i: C.swiglu_bwd(act.data_ptr(), gu.data_ptr(), dgu.data_ptr(), T, inter, s))
    b = 5 * T * inter * 2
    res["elementwise"].append({"name": "swiglu_bwd", "ms": ms, "gbs": b / ms / 1e6, "frac_of_hbm": b / ms / 1e6 / hbm})
    n = 1 << 30
    p = torch.zeros(n, device="cuda", dtype=torch.bfloat16)
    g = torch.zeros(n, device="cuda", dtype=torch.bfloat16)
    master = torch.zeros(n, device="cuda")
    for sd, dt, nb in (("bf16", torch.bfloat16, 2), ("fp32", torch.float32, 4)):
        m = torch.zeros(n, device="cuda", dtype=dt)
        v = torch.zeros(n, device="cuda", dtype=dt)
        ms = time_fn(lambda i: C.adamw_flat(p.data_ptr(), g.data_ptr(), master.data_ptr(), m.data_ptr(), v.data_ptr(),
                                            int(sd == "bf16"), n, 1e-4, 0.9, 0.999, 1e-8, 0.01, 0.1, 0.001, 1.0, 7, s),
                     iters=5, warm=2)
        b = n * (2 + 2 + 8 + 4 * nb)
        res["elementwise"].append({"name": f"adamw_flat[{sd} moments] n=2^30", "ms": ms, "gbs": b / ms / 1e6,
                                   "frac_of_hbm": b / ms / 1e6 / hbm})
        del m, v
    theta = torch.zeros(n, device="cuda")
    mom = torch.zeros(n, device="cuda")
    ms = time_fn(lambda i: C.diloco_outer_local(master.data_ptr(), 0, 0, theta.data_ptr(), mom.data_ptr(),
                                                master.data_ptr(), p.data_ptr(), n, 0.7, 0.9, 0, 1.0, s), iters=5, warm=2)
    b = n * (4 + 8 + 8 + 4 + 2)
    res["elementwise"].append({"name": "diloco_outer_local n=2^30", "ms": ms, "gbs": b / ms / 1e6,
                               "frac_of_hbm": b / ms / 1e6 / hbm})
    for e in res["elementwise"]:
        print(json.dumps(e), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/kernel_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
