"""Survivors of a fused outer step settle on one verdict before anyone retries (per-rank barrier time-outs
can disagree; a rank that passed barrier 1 has already changed its state)."""
import pytest
import torch

from hypha_b200.parallel.diloco import DiLoCoOuter


def test_merge_verdicts():
    m = DiLoCoOuter.merge_verdicts
    assert m([]) == (0, 0)
    assert m([(0, 0), (0, 0)]) == (0, 0)
    assert m([(1, 0b0100), (1, 0b0100), (1, 0b1100)]) == (1, 0b1100)       # nobody passed: union of the missing
    # rank A saw the straggler arrive inside its window, passed barrier 1 and ran phase B; B timed out: not retryable
    assert m([(0, 0), (1, 0b0100)])[0] == 2
    assert m([(2, 0b0100), (1, 0b0100)])[0] == 2
    assert m([(2, 0)])[0] == 2


def _outer(world, rank, script):
    """A DiLoCoOuter whose launches are scripted: script = [(status, missing), ...] per attempt."""
    o = object.__new__(DiLoCoOuter)
    o.world, o.rank, o.round, o.backend = world, rank, 3, "fused"
    o.n, o.segments = 1024 * world, {rank: ("own",)}
    o.set_members(list(range(world)))
    attempts = iter(script)
    o.step = lambda: setattr(o, "round", o.round + 1)
    o.check = lambda: next(attempts)
    dropped = []
    o.drop_ranks = lambda dead: (dropped.append(list(dead)), o.set_members([r for r in o.members() if r not in dead]))
    return o, dropped


def test_step_resilient_agrees_before_retrying():
    # ranks 0 and 1 of 4; rank 2 is dead, rank 3 is slow: rank 0 saw only 2 missing, rank 1 saw 2 and 3
    o0, d0 = _outer(4, 0, [(1, 0b0100), (0, 0)])
    agreed = []

    def agree0(v):
        agreed.append(v)
        return [v, (1, 0b1100)] if v[0] else [v, (0, 0)]

    assert o0.step_resilient(agree=agree0) == (0, 0)
    assert d0 == [[2, 3]] and o0.members() == [0, 1] and o0.round == 4 and agreed[0] == (1, 0b0100)
    # the dangerous interleaving: this rank timed out, a peer went through -> everybody stops, nothing is dropped
    o1, d1 = _outer(4, 1, [(1, 0b0100)])
    with pytest.raises(RuntimeError, match="mid-broadcast"):
        o1.step_resilient(agree=lambda v: [v, (0, 0)])
    assert d1 == [] and o1.round == 3
    # without a control plane every rank acts on its own status (single-process drills)
    o2, d2 = _outer(2, 0, [(1, 0b10), (0, 0)])
    assert o2.step_resilient() == (0, 0) and d2 == [[1]]


def test_outer_state_is_bound_to_its_shard_and_team():
    o = object.__new__(DiLoCoOuter)
    o.world, o.rank, o.backend, o.round = 4, 1, "nccl", 0
    o.shard_begin, o.shard_len = 256, 256
    o.theta_outer, o.momentum = torch.zeros(256), torch.zeros(256)
    good = {"theta_outer": torch.ones(256), "momentum": torch.ones(256), "round": torch.tensor(7),
            "shard_begin": torch.tensor(256), "world": torch.tensor(4)}
    o.load_state_dict(good)
    assert o.round == 7 and float(o.theta_outer[0]) == 1.0
    with pytest.raises(ValueError, match="team of 8"):
        o.load_state_dict(dict(good, world=torch.tensor(8)))
    with pytest.raises(ValueError, match="shard at 0"):
        o.load_state_dict(dict(good, shard_begin=torch.tensor(0)))
    with pytest.raises(ValueError, match="elements"):
        o.load_state_dict(dict(good, theta_outer=torch.ones(128)))
