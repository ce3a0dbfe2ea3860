"""A replacement PROCESS joins the fused outer step (experimental: `DiLoCoOuter.admit`, `PeerGroup.rebind`).

Not run under torchrun: this script is its own launcher because the newcomer must be a process that was not
part of the original group.

    python tests/mgpu_join_check.py [--world 3]

  parent        picks two free ports, starts `world` rank processes, waits for the victim to exit, then
                starts the replacement process
  ranks         gloo control group -> PeerGroup -> fused DiLoCoOuter; round 0 with everyone; the last rank
                then exits hard (os._exit) without releasing anything
  survivors     drop it (control-plane knowledge, as after a lost lease), run round 1 alone, then `admit` on a
                TCPStore the parent announced
  replacement   builds the same buffers on `PeerGroup.joining(dead, world, dev)`, `admit(joining=True)`:
                maps everyone, pulls the weights from rank 0
  everyone      round 2 with the full membership -> identical parameters on all ranks
"""
import argparse
import json
import os
import socket
import subprocess
import sys
import time
from datetime import timedelta

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
N = 8 * 1024 * 1024


def free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def build(grp, dev, seed_params=True):
    import torch

    from hypha_b200.parallel.diloco import AdamWConfig, DiLoCoConfig, DiLoCoOuter, FlatAdamW

    pbuf = grp.alloc("model.params", N, torch.bfloat16)
    gbuf = grp.alloc("model.grads", N, torch.bfloat16)
    if seed_params:
        g0 = torch.Generator(device=dev).manual_seed(7)
        pbuf.local.copy_(torch.randn(N, device=dev, generator=g0))
    opt = FlatAdamW(pbuf.local, gbuf.local, AdamWConfig(state_dtype="fp32"))
    outer = DiLoCoOuter(opt, DiLoCoConfig(backend="fused", outer_lr=0.7, outer_momentum=0.9, timeout_ms=3000),
                        grp, pbuf, gbuf)
    return pbuf, opt, outer


def fingerprint(pbuf):
    import torch

    torch.cuda.synchronize()
    return (float(pbuf.local.float().sum().item()), float(pbuf.local.float().abs().sum().item()))


def rank_main(rank: int, world: int, gloo_port: int, store_port: int, joining: bool) -> int:
    import torch
    import torch.distributed as dist

    from hypha_b200.parallel.peer_memory import PeerGroup, store_exchange

    os.environ["HYPHA_NVLS"] = "0"
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dead = world - 1
    order = list(range(world))
    gr = torch.Generator(device=dev).manual_seed(100 + rank + (1000 if joining else 0))

    def admit_exchange():
        store = dist.TCPStore("127.0.0.1", store_port, world, rank == 0, timedelta(seconds=180), wait_for_workers=False)
        return store, store_exchange(store, rank, world, prefix="gen1")

    if joining:
        grp = PeerGroup.joining(dead, world, rank)
        pbuf, opt, outer = build(grp, dev, seed_params=False)
        store, ex = admit_exchange()
        outer.admit(ex, order, [dead], True)
    else:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{gloo_port}", rank=rank, world_size=world)
        grp = PeerGroup.from_torch_distributed(rank)
        pbuf, opt, outer = build(grp, dev)
        opt.master.add_(0.01 * torch.randn(N, device=dev, generator=gr))
        assert outer.step_resilient() == (0, 0)                       # round 0: everyone
        fps = [None] * world
        dist.all_gather_object(fps, fingerprint(pbuf))
        assert all(f == fps[0] for f in fps), fps
        if rank == dead:
            print("victim exits", flush=True)
            os._exit(0)                                               # no cleanup: mappings of it go stale
        outer.drop_ranks([dead])                                      # the scheduler said so (lost lease)
        opt.master.add_(0.01 * torch.randn(N, device=dev, generator=gr))
        assert outer.step_resilient() == (0, 0)                       # round 1: survivors
        store, ex = admit_exchange()
        outer.admit(ex, order, [dead], False)
    assert outer.members() == order and (rank == dead or dead not in outer.segments)
    fps = ex(fingerprint(pbuf))
    assert all(f == fps[0] for f in fps), ("after admission", fps)
    opt.master.add_(0.01 * torch.randn(N, device=dev, generator=gr))
    assert outer.step_resilient() == (0, 0)                           # round 2: full membership again
    fps = ex(fingerprint(pbuf))
    assert all(f == fps[0] for f in fps), ("round 2", fps)
    pulled = grp.counters.snapshot()["hypha.bandwidth.inbound.bytes"].get("/hypha-peer-memory/pull", 0)
    assert pulled == (2 * N if joining else 0)
    ex("done")
    if rank == 0:
        print(json.dumps({"world": world, "replaced_rank": dead, "correct": True}), flush=True)
    os._exit(0)      # the gloo group still contains the dead process: no orderly teardown


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=3)
    ap.add_argument("--rank", type=int, default=-1)
    ap.add_argument("--gloo-port", type=int, default=0)
    ap.add_argument("--store-port", type=int, default=0)
    ap.add_argument("--joining", action="store_true")
    a = ap.parse_args()
    if a.rank >= 0:
        return rank_main(a.rank, a.world, a.gloo_port, a.store_port, a.joining)
    gp, sp = free_port(), free_port()
    base = [sys.executable, os.path.abspath(__file__), "--world", str(a.world), "--gloo-port", str(gp), "--store-port", str(sp)]
    procs = [subprocess.Popen(base + ["--rank", str(r)]) for r in range(a.world)]
    victim = procs[a.world - 1]
    try:
        victim.wait(timeout=240)
        time.sleep(1.0)
        procs.append(subprocess.Popen(base + ["--rank", str(a.world - 1), "--joining"]))
        rc = 0
        for p in procs:
            rc |= p.wait(timeout=300)
        return rc
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()


if __name__ == "__main__":
    sys.exit(main())
