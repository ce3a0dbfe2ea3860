"""Fault-tolerance check of the fused outer step (BASELINE config 4 at kernel level), run under
torchrun with >= 2 ranks.  The control plane here is a *gloo* process group (handle exchange,
round-boundary coordination): no NCCL anywhere -- the data path is the peer-memory kernel only.

  round 0   all ranks                       -> normal fused outer step
  round 1   last rank stalls ("killed")     -> survivors time out at barrier 1, drop it, adopt its
                                               shard (theta from the bf16 anchor), retry, succeed
  round 2   survivors only
  rejoin    the stalled rank pulls the current weights from rank 0 over peer memory
            (parameter broadcast), everyone re-admits it and re-syncs the flag epoch
  round 3   all ranks again                 -> identical parameters everywhere
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypha_b200.parallel.diloco import AdamWConfig, DiLoCoConfig, DiLoCoOuter, FlatAdamW  # noqa: E402
from hypha_b200.parallel.peer_memory import PeerGroup  # noqa: E402


def gather(obj, world):
    out = [None] * world
    dist.all_gather_object(out, obj)
    return out


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("gloo")
    grp = PeerGroup.from_torch_distributed(local)
    n = 8 * 1024 * 1024
    pbuf = grp.alloc("model.params", n, torch.bfloat16)
    gbuf = grp.alloc("model.grads", n, torch.bfloat16)
    g0 = torch.Generator(device=dev).manual_seed(7)
    pbuf.local.copy_(torch.randn(n, device=dev, generator=g0))
    opt = FlatAdamW(pbuf.local, gbuf.local, AdamWConfig(state_dtype="fp32"))
    outer = DiLoCoOuter(opt, DiLoCoConfig(backend="fused", outer_lr=0.7, outer_momentum=0.9, timeout_ms=1500),
                        grp, pbuf, gbuf)
    gr = torch.Generator(device=dev).manual_seed(100 + rank)
    dead = world - 1
    log = []

    def drift():
        opt.master.add_(0.01 * torch.randn(n, device=dev, generator=gr))

    def fingerprint():
        torch.cuda.synchronize()
        return (float(pbuf.local.float().sum().item()), float(pbuf.local.float().abs().sum().item()))

    # ---- round 0: everyone ----------------------------------------------------------------
    drift()
    assert outer.step_resilient() == (0, 0)
    fps = gather(fingerprint(), world)
    assert all(f == fps[0] for f in fps), fps
    log.append("round0 ok")

    # ---- round 1: the last rank stalls -------------------------------------------------------
    drift()
    t0 = time.time()
    if rank != dead:
        assert outer.step_resilient() == (0, 0)
        assert outer.members() == [r for r in range(world) if r != dead]
        took = time.time() - t0
        assert took >= 1.4, f"timeout path not exercised ({took:.2f}s)"
        if outer.adopter_of(dead, outer.members()) == rank:
            assert dead in outer.segments
    fps = gather(fingerprint() if rank != dead else None, world)
    live = [f for f in fps if f is not None]
    assert all(f == live[0] for f in live), fps
    log.append("round1 (one rank missing) ok")

    # ---- round 2: survivors only ----------------------------------------------------------------
    drift()
    if rank != dead:
        assert outer.step_resilient() == (0, 0)
    fps = gather(fingerprint() if rank != dead else None, world)
    live = [f for f in fps if f is not None]
    assert all(f == live[0] for f in live), fps
    log.append("round2 (survivors) ok")

    # ---- rejoin: parameter broadcast over peer memory ---------------------------------------------
    dist.barrier()
    if rank == dead:
        outer.reseed_from(0)
        torch.cuda.synchronize()
    dist.barrier()
    # the holder of the dead rank's shard kept its fp32 theta_outer / momentum going for two rounds: `readmit`
    # must hand exactly that state back (no re-derivation from bf16, no momentum reset)
    held = None
    if rank != dead and dead in outer.segments:
        _, _, th, mo = outer.segments[dead]
        torch.cuda.synchronize()
        held = (float(th.double().sum().item()), float(mo.double().sum().item()), float(mo.abs().max().item()))
    outer.readmit(dead)
    outer.sync_epoch()
    torch.cuda.synchronize()
    mine = (float(outer.theta_outer.double().sum().item()), float(outer.momentum.double().sum().item()),
            float(outer.momentum.abs().max().item())) if rank == dead else None
    helds, mines = gather(held, world), gather(mine, world)
    want = [h for h in helds if h is not None]
    assert len(want) == 1 and want[0] == mines[dead], ("hand-back mismatch", want, mines[dead])
    assert want[0][2] > 0.0, "the adopted shard's momentum should be non-zero after two survivor rounds"
    log.append("hand-back of theta_outer + momentum to the returning owner ok")
    fps = gather(fingerprint(), world)
    assert all(f == fps[0] for f in fps), ("after rejoin", fps)
    assert outer.members() == list(range(world)) and (rank == dead or dead not in outer.segments)
    log.append("rejoin (peer-memory parameter broadcast) ok")

    # ---- round 3: everyone again -------------------------------------------------------------------
    drift()
    assert outer.step_resilient() == (0, 0)
    fps = gather(fingerprint(), world)
    assert all(f == fps[0] for f in fps), fps
    log.append("round3 (full membership) ok")
    pulled = grp.counters.snapshot()["hypha.bandwidth.inbound.bytes"].get("/hypha-peer-memory/pull", 0)
    if rank == dead:
        assert pulled == 2 * n
        hb = grp.counters.snapshot()["hypha.bandwidth.inbound.bytes"].get("/hypha-peer-memory/handback", 0)
        assert hb == 2 * 4 * outer.shard_len, hb
    if rank == 0:
        print(json.dumps({"world": world, "steps": log, "correct": True}), flush=True)
    dist.barrier()
    grp.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
