"""Worker of `test_sharded_outer.py`: one rank of a gloo team running `DiLoCoOuter(backend="sharded")` (the outer step
of a team that spans more than one NVSwitch box: reduce-scatter / all-gather of the process group, fp32 outer state
sharded over all ranks) against a single-process model of the same arithmetic.

    RANK=r WORLD_SIZE=w MASTER_ADDR=127.0.0.1 MASTER_PORT=p python tests/mp_sharded_outer_check.py <n>
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from hypha_b200.parallel.diloco import AdamWConfig, DiLoCoConfig, DiLoCoOuter, FlatAdamW  # noqa: E402

LR, MU, ROUNDS = 0.7, 0.9, 3


def progress(rank: int, rnd: int, n: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(1000 * rnd + rank)
    return torch.randn(n, generator=g) * 1e-2


def main() -> None:
    n = int(sys.argv[1])
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    theta0 = (torch.randn(n, generator=torch.Generator().manual_seed(7)) * 0.05).to(torch.bfloat16)
    params = theta0.clone()
    opt = FlatAdamW(params, torch.zeros_like(params), AdamWConfig())
    outer = DiLoCoOuter(opt, DiLoCoConfig(backend="sharded", outer_lr=LR, outer_momentum=MU))
    assert outer.backend == "sharded" and (outer.world, outer.rank) == (world, rank)
    lens = [None] * world
    dist.all_gather_object(lens, (outer.shard_begin, outer.shard_len))
    assert sum(l for _, l in lens) == n and all(lens[i][0] + lens[i][1] == lens[i + 1][0] or lens[i + 1][1] == 0
                                                for i in range(world - 1)), lens
    assert outer.theta_outer.numel() == outer.shard_len          # 1 / world of the fp32 state, not a replica

    # the model: full-length fp32 theta_outer / momentum, every rank's local weights
    theta, mom, anchor = theta0.float(), torch.zeros(n), theta0.clone()
    saved = None
    for rnd in range(ROUNDS):
        local = [anchor.float() + progress(r, rnd, n) for r in range(world)]       # what H inner steps produced
        opt.master.copy_(local[rank])
        opt.params.copy_(opt.master)
        if rnd == 1:                                             # a checkpoint in the middle of a round
            saved = ({k: v.clone() for k, v in outer.state_dict().items()}, opt.master.clone())
        outer.step()
        g = sum(l - anchor.float() for l in local) / world
        mom = g.clone() if rnd == 0 else MU * mom + g
        theta = theta + LR * (MU * mom + g)
        anchor = theta.to(torch.bfloat16)
        # the order of the fp32 sum over ranks is the collective's: a value on a rounding tie may land on the
        # neighbouring bf16 number
        off = opt.params != anchor
        assert off.float().mean() < 1e-3 and (opt.params.float() - anchor.float()).abs().max() < 1e-3, rnd
        anchor = opt.params.clone()
        assert torch.equal(opt.master, anchor.float())           # every replica restarts from the same bf16 weights
        sl = slice(outer.shard_begin, outer.shard_begin + outer.shard_len)
        torch.testing.assert_close(outer.theta_outer, theta[sl], rtol=0, atol=1e-6)
        torch.testing.assert_close(outer.momentum, mom[sl], rtol=0, atol=1e-6)
        everyone = [torch.empty_like(opt.params) for _ in range(world)]
        dist.all_gather(everyone, opt.params)
        assert all(torch.equal(e, opt.params) for e in everyone)
    assert outer.round == ROUNDS

    # resume: a fresh optimizer pair restored from the mid-run checkpoint repeats rounds 1.. bit for bit
    sd, master = saved
    p2 = theta0.clone()
    opt2 = FlatAdamW(p2, torch.zeros_like(p2), AdamWConfig())
    outer2 = DiLoCoOuter(opt2, DiLoCoConfig(backend="sharded", outer_lr=LR, outer_momentum=MU))
    outer2.load_state_dict(sd)
    opt2.master.copy_(master)
    assert outer2.round == 1
    for rnd in range(1, ROUNDS):
        if rnd > 1:
            opt2.master.copy_(outer2.anchor.float() + progress(rank, rnd, n))
        outer2.step()
    assert torch.equal(opt2.params, opt.params) and torch.equal(outer2.theta_outer, outer.theta_outer)
    bad = dict(sd, world=torch.tensor(world + 1))
    try:
        outer2.load_state_dict(bad)
        raise AssertionError("a checkpoint of another team size was accepted")
    except ValueError:
        pass
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank}: sharded outer step ok (n={n}, shard {outer.shard_len})", flush=True)


if __name__ == "__main__":
    main()
