#!/usr/bin/env python
"""Outer-step bandwidth sweep (BASELINE config 5): fused peer-memory DiLoCo outer step vs NCCL.

    torchrun --nproc-per-node N bench/outer_sweep.py [--min-mb 1] [--max-mb 16384] --json out.json

For each bf16 delta size S (1 MB .. 16 GB) it times, on the device with CUDA events (max over
ranks, warm-up first):
  * fused     -- ONE kernel: delta, cross-rank mean over P2P loads, Nesterov, P2P broadcast,
                 master/anchor refresh (hypha_b200/csrc/kernels/outer_step.cu)
  * nccl_ar   -- `dist.all_reduce` of the same bf16 delta only (collective cost alone)
  * nccl_full -- the reference-semantics baseline: delta + all_reduce + unfused torch Nesterov +
                 parameter refresh (chunked so it fits in memory)
Bus bandwidth uses the nccl-tests convention  S * 2(N-1)/N / t.  Roofline denominators: the
measured 770 GB/s peer copy (B200_PROFILING.md) and the nominal 900 GB/s NVLink 5 direction."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypha_b200.parallel.diloco import DiLoCoConfig, DiLoCoOuter  # noqa: E402
from hypha_b200.parallel.peer_memory import PeerGroup  # noqa: E402

PEER_COPY_GBS, NVLINK_NOMINAL_GBS = 770.0, 900.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min-mb", type=float, default=1)
    ap.add_argument("--max-mb", type=float, default=16384)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--json", default=None)
    ap.add_argument("--nvls", action="store_true", help="also time the NVSwitch-multicast (multimem) variant")
    ap.add_argument("--tune-mb", type=float, default=0, help="at this size also sweep chunks x link CTAs")
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    grp = PeerGroup.from_torch_distributed(local)
    nmax = int(a.max_mb * 1024 * 1024 / 2)
    nmax = (nmax + 8191) // 8192 * 8192
    if a.nvls:
        from hypha_b200.parallel.nvls import alloc_nvls

        pbuf = alloc_nvls(grp, "params", nmax, torch.bfloat16)
        gbuf = alloc_nvls(grp, "grads", nmax, torch.bfloat16)
    else:
        pbuf = grp.alloc("params", nmax, torch.bfloat16)
        gbuf = grp.alloc("grads", nmax, torch.bfloat16)
    master_full = torch.zeros(nmax, device=dev)

    def timed(fn, iters):
        fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    rows = []
    mb = a.min_mb
    while mb <= a.max_mb * 1.0001:
        n = int(mb * 1024 * 1024 / 2)
        n = max(8192, (n + 8191) // 8192 * 8192)
        p, g = pbuf.local[:n], gbuf.local[:n]

        class _Opt:                      # the pieces of FlatAdamW the outer step touches
            params, grads, master = p, g, master_full[:n]

        kw = (lambda b: {"mc_ptr": b.mc_ptr}) if a.nvls else (lambda b: {})
        sub_p = type(pbuf)(pbuf.name, n, pbuf.dtype, p, pbuf.peer_ptrs, **kw(pbuf))
        sub_g = type(gbuf)(gbuf.name, n, gbuf.dtype, g, gbuf.peer_ptrs, **kw(gbuf))
        outer = DiLoCoOuter(_Opt, DiLoCoConfig(backend="fused"), grp, sub_p, sub_g)
        iters = a.iters if mb >= 64 else 20
        st = (0, 0)

        def variant(nvls, chunks, link_ctas=0):
            nonlocal st
            outer.use_nvls = nvls
            if chunks == 0 and world <= 2:
                chunks = max(1, min(16, (2 * n) // (128 << 20)))     # force the pipeline for the comparison
            os.environ["HYPHA_OUTER_CHUNKS"] = str(chunks)
            os.environ["HYPHA_OUTER_LINK_CTAS"] = str(link_ctas)
            ms = timed(outer.step, iters)
            st = tuple(max(x, y) for x, y in zip(st, outer.check()))
            return ms

        plain = variant(False, 1)                       # A | B | C, P2P loads/stores
        piped = variant(False, 0)                       # auto-chunked software pipeline
        fused = min(plain, piped)
        nvls_ms = nvls_plain = nvls_piped = None
        if a.nvls:
            nvls_plain = variant(True, 1)
            nvls_piped = variant(True, 0)
            nvls_ms = min(nvls_plain, nvls_piped)
        tune = None
        if a.tune_mb and abs(mb - a.tune_mb) < 1e-6:
            tune = []
            for nv in ([False, True] if a.nvls else [False]):
                for k in (2, 4, 8, 16):
                    for b in (16, 32, 48, 72, 100):
                        tune.append({"nvls": nv, "chunks": k, "link_ctas": b, "ms": variant(nv, k, b)})
            if rank == 0:
                best = min(tune, key=lambda t: t["ms"])
                print("tune best:", json.dumps(best), flush=True)
            os.environ["HYPHA_OUTER_CHUNKS"] = os.environ["HYPHA_OUTER_LINK_CTAS"] = "0"
        ar = timed(lambda: dist.all_reduce(g), iters)
        CH = 1 << 28
        try:
            th, mo = _Opt.master.clone(), torch.zeros(n, device=dev)
        except torch.OutOfMemoryError:
            th = mo = None
            torch.cuda.empty_cache()

        def nccl_full():
            d = torch.empty(n, device=dev, dtype=torch.bfloat16)
            for s in range(0, n, CH):
                e = min(n, s + CH)
                d[s:e] = (_Opt.master[s:e] - th[s:e]).to(torch.bfloat16)
            dist.all_reduce(d)
            for s in range(0, n, CH):
                e = min(n, s + CH)
                gg = d[s:e].float().div_(world)
                mo[s:e].mul_(0.9).add_(gg)
                th[s:e].add_((mo[s:e] * 0.9).add_(gg).mul_(0.7))
                _Opt.master[s:e].copy_(th[s:e])
                p[s:e].copy_(th[s:e])

        try:
            full = timed(nccl_full, max(2, iters // 2)) if th is not None else None
        except torch.OutOfMemoryError:      # the unfused baseline needs ~10 more bytes/element
            full = None
            torch.cuda.empty_cache()
        del th, mo
        S = 2 * n
        bus = lambda ms: S * 2 * (world - 1) / world / ms / 1e6
        row = {"delta_mb": S / 2 ** 20, "numel": n, "status": list(st), "fused_ms": fused, "nccl_allreduce_ms": ar,
               "nccl_full_outer_ms": full, "fused_busbw_gbs": bus(fused),
               "fused_plain_ms": plain, "fused_pipelined_ms": piped, "pipeline_shape": list(outer._pipeline_shape()),
               "fused_nvls_ms": nvls_ms, "fused_nvls_busbw_gbs": bus(nvls_ms) if nvls_ms else None,
               "nvls_plain_ms": nvls_plain, "nvls_pipelined_ms": nvls_piped, "tune": tune, "nccl_allreduce_busbw_gbs": bus(ar),
               "fused_frac_of_770": bus(fused) / PEER_COPY_GBS, "fused_frac_of_900": bus(fused) / NVLINK_NOMINAL_GBS,
               "speedup_vs_nccl_full": (full / fused) if full else None}
        rows.append(row)
        if rank == 0:
            print(json.dumps(row), flush=True)
        # release this size's anchor/flag buffers
        del outer
        mb *= 4
    if rank == 0 and a.json:
        os.makedirs(os.path.dirname(a.json) or ".", exist_ok=True)
        json.dump({"world": world, "rows": rows}, open(a.json, "w"), indent=1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
