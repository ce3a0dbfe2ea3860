"""Multi-GPU check + micro-benchmark of the fused peer-memory DiLoCo outer step (run under
torchrun, one rank per GPU).  Verifies against a reference built from NCCL all-reduces of the
same bf16 deltas, then times fused vs NCCL-baseline outer steps (CUDA events, max over ranks).

    torchrun --nproc-per-node 2 tests/mgpu_outer_check.py [--numel N] [--json out.json]
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypha_b200.parallel.diloco import AdamWConfig, DiLoCoConfig, DiLoCoOuter, FlatAdamW  # noqa: E402
from hypha_b200.parallel.peer_memory import PeerGroup  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--numel", type=int, default=64 * 1024 * 1024)
    ap.add_argument("--json", default=None)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--chunks", type=int, default=0, help="software-pipeline depth of the fused kernel (0 = auto)")
    ap.add_argument("--nvls", action="store_true", help="NVSwitch multicast buffers (multimem path)")
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    grp = PeerGroup.from_torch_distributed(local)
    n = a.numel
    if a.nvls:
        from hypha_b200.parallel.nvls import alloc_nvls, multicast_supported

        if not multicast_supported(local):
            if rank == 0:
                print(json.dumps({"correct": True, "skipped": "multicast not supported on this device"}), flush=True)
            dist.destroy_process_group()
            sys.exit(0)
        pbuf = alloc_nvls(grp, "params", n, torch.bfloat16)
        gbuf = alloc_nvls(grp, "grads", n, torch.bfloat16)
        assert pbuf.mc_ptr and gbuf.mc_ptr
    else:
        pbuf = grp.alloc("params", n, torch.bfloat16)
        gbuf = grp.alloc("grads", n, torch.bfloat16)
    g = torch.Generator(device=dev).manual_seed(1)
    pbuf.local.copy_(torch.randn(n, device=dev, generator=g))          # identical start on every rank
    opt = FlatAdamW(pbuf.local, gbuf.local, AdamWConfig())
    cfg = DiLoCoConfig(backend="fused", outer_lr=0.7, outer_momentum=0.9, pipeline_chunks=a.chunks, link_ctas=40)
    outer = DiLoCoOuter(opt, cfg, grp, pbuf, gbuf)
    ok = True
    gr = torch.Generator(device=dev).manual_seed(100 + rank)
    th_full = opt.master.clone()                                       # replicated fp32 reference state
    mom_full = torch.zeros(n, device=dev)
    for rnd in range(3):
        opt.master.add_(0.01 * torch.randn(n, device=dev, generator=gr))   # "H inner steps"
        anchor = outer.anchor_buf.local.float()
        delta = (opt.master - anchor).to(torch.bfloat16).float()
        dist.all_reduce(delta)
        if a.nvls:
            # the switch accumulates in fp32 but returns the N-way sum as bf16 (the precision of a bf16
            # NCCL all-reduce); the P2P path keeps the fp32 sum
            delta = delta.to(torch.bfloat16).float()
        gmean = delta / world + (anchor - th_full)
        mom_full = gmean.clone() if rnd == 0 else 0.9 * mom_full + gmean
        th_full = th_full + 0.7 * (0.9 * mom_full + gmean)
        outer.step()
        st = outer.check()
        dist.barrier()
        b, l = outer.shard_begin, outer.shard_len
        e1 = (outer.theta_outer - th_full[b:b + l]).abs().max().item()
        want_p = th_full.to(torch.bfloat16)
        mism = (pbuf.local != want_p).float().mean().item()           # bf16 ties may round differently
        e3 = (opt.master - pbuf.local.float()).abs().max().item()
        e4 = (outer.anchor_buf.local != pbuf.local).sum().item()
        # NVLS: the in-switch reduction returns a bf16 sum (rounding mode unspecified) -> 1 bf16 ulp of
        # the largest |sum| (~0.1) times lr*(1+mu)/N; replicas must still agree bit-for-bit (checked below)
        tol_t, tol_p = (1e-3, 5e-2) if a.nvls else (1e-4, 1e-3)
        good = st == (0, 0) and e1 < tol_t and mism < tol_p and e3 == 0 and e4 == 0
        ok = ok and good
        if rank == 0:
            print(f"round {rnd}: status={st} |theta-ref|max={e1:.2e} param-mismatch-frac={mism:.2e} "
                  f"master==param:{e3 == 0} anchor==param:{e4 == 0} -> {'OK' if good else 'FAIL'}", flush=True)
    # every rank must hold identical parameters
    ref = pbuf.local.clone()
    dist.broadcast(ref, 0)
    same = bool((ref == pbuf.local).all().item())
    ok = ok and same

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            fn()
        e1_.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1_) / a.iters], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    fused_ms = timed(outer.step)
    # NCCL baseline: bf16 all_reduce of the delta + unfused torch Nesterov (chunked to bound memory)
    th_b, mom_b = opt.master.clone(), torch.zeros(n, device=dev)

    def nccl_step():
        d = (opt.master - th_b).to(torch.bfloat16)
        dist.all_reduce(d)
        gg = d.float().div_(world)
        mom_b.mul_(0.9).add_(gg)
        th_b.add_((mom_b * 0.9).add_(gg).mul_(0.7))
        opt.master.copy_(th_b)
        pbuf.local.copy_(th_b)

    nccl_ms = timed(nccl_step)

    def allreduce_only():
        dist.all_reduce(gbuf.local)

    ar_ms = timed(allreduce_only)
    bytes_ = 2 * n
    busbw = lambda ms: bytes_ * 2 * (world - 1) / world / ms / 1e6     # GB/s, nccl-tests convention
    res = {"world": world, "numel": n, "nvls": bool(a.nvls and outer._multicast_ptrs()[0]), "delta_bytes": bytes_, "correct": bool(ok), "identical_params": same,
           "fused_outer_ms": fused_ms, "nccl_outer_ms": nccl_ms, "nccl_allreduce_only_ms": ar_ms,
           "fused_busbw_gbs": busbw(fused_ms), "nccl_allreduce_busbw_gbs": busbw(ar_ms),
           "speedup_vs_nccl_outer": nccl_ms / fused_ms}
    if rank == 0:
        print(json.dumps(res), flush=True)
        if a.json:
            os.makedirs(os.path.dirname(a.json) or ".", exist_ok=True)
            json.dump(res, open(a.json, "w"), indent=1)
    dist.barrier()
    grp.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
