"""`PeerGroup.rebind` / `PeerGroup.joining` / `store_exchange`: re-mapping symmetric buffers when a PROCESS of
the group is replaced.  The CUDA-IPC calls are faked (pointers are integers, a mapping remembers what it
maps), ranks are threads; the store exchange runs over a real TCPStore."""
import itertools
import threading
import types

import pytest
import torch

from hypha_b200.parallel import peer_memory as pm


class FakeC:
    def __init__(self):
        self.ids = itertools.count(0x1000, 0x1000)
        self.lock = threading.Lock()
        self.allocs = {}          # ptr -> alive
        self.maps = {}            # mapping ptr -> target ptr
        self.closed = []

    def sym_alloc(self, nbytes, zero):
        with self.lock:
            p = next(self.ids)
            self.allocs[p] = True
            return p

    def sym_free(self, p):
        self.allocs[p] = False

    def ipc_export(self, ptr):
        assert self.allocs.get(ptr), "exporting memory that is not a live allocation"
        return f"h{ptr}".encode()

    def ipc_open(self, handle):
        target = int(handle[1:])
        if not self.allocs.get(target):
            raise RuntimeError("cudaIpcOpenMemHandle: invalid handle (exporter is gone)")
        with self.lock:
            p = next(self.ids)
            self.maps[p] = target
            return p

    def ipc_close(self, p):
        self.closed.append(p)
        self.maps.pop(p)


def _barrier_exchange(world):
    shared = {"slots": [None] * world, "barrier": threading.Barrier(world)}

    def make(rank):
        def exchange(obj):
            shared["slots"][rank] = obj
            shared["barrier"].wait(30)
            out = list(shared["slots"])
            shared["barrier"].wait(30)
            return out
        return exchange
    return make


def _in_threads(fns):
    res = [None] * len(fns)

    def run(i):
        try:
            res[i] = fns[i]()
        except Exception as exc:      # noqa: BLE001
            res[i] = exc
    ts = [threading.Thread(target=run, args=(i,)) for i in range(len(fns))]
    for t in ts:
        t.start()
    for t in ts:
        t.join(60)
    assert not any(t.is_alive() for t in ts)
    return res


@pytest.fixture
def fake(monkeypatch):
    C = FakeC()
    monkeypatch.setattr(pm, "load_C", lambda: C)
    monkeypatch.setattr(pm, "tensor_from_ptr", lambda ptr, numel, dtype, device: types.SimpleNamespace(
        data_ptr=lambda: ptr, numel=lambda: numel))

    class _Dev:
        def __init__(self, d): pass
        def __enter__(self): return self
        def __exit__(self, *a): return False
    monkeypatch.setattr(pm, "torch", types.SimpleNamespace(cuda=types.SimpleNamespace(device=_Dev), dtype=torch.dtype,
                                                           bfloat16=torch.bfloat16, float32=torch.float32))
    return C


def _group_of(world, C, shapes):
    make = _barrier_exchange(world)
    groups = [pm.PeerGroup(r, world, r, make(r)) for r in range(world)]

    def alloc(g):
        for name, (n, dt) in shapes.items():
            g.alloc(name, n, dt)
        return True
    assert all(r is True for r in _in_threads([lambda g=g: alloc(g) for g in groups]))
    return groups


SHAPES = {"model.params": (1 << 20, torch.bfloat16), "outer.flags": (64, torch.float32)}


def _resolve(C, g, name, r):
    p = g.buffers[name].peer_ptrs[r]
    return p if r == g.rank else C.maps[p]


def test_replacement_process_is_mapped_by_everyone(fake):
    C, world, dead = fake, 4, 2
    groups = _group_of(world, C, SHAPES)
    for g in groups:                                   # sanity: everybody sees everybody's local buffer
        for name in SHAPES:
            assert [_resolve(C, g, name, r) for r in range(world)] == [x.buffers[name].local.data_ptr() for x in groups]
    before = {(g.rank, name): list(g.buffers[name].peer_ptrs) for g in groups for name in SHAPES}
    # the process of rank 2 dies: its allocations are gone, mappings OF it are stale
    for name in SHAPES:
        C.allocs[groups[dead].buffers[name].local.data_ptr()] = False
    new = pm.PeerGroup.joining(dead, world, dead)
    for name, (n, dt) in SHAPES.items():
        new.alloc(name, n, dt)                         # no exchange: nobody else is allocating
        assert new.buffers[name].peer_ptrs.count(0) == world - 1
    members = [g for g in groups if g.rank != dead] + [new]
    members.sort(key=lambda g: g.rank)
    make = _barrier_exchange(world)
    res = _in_threads([lambda g=g: g.rebind(make(g.rank), [dead]) for g in members])
    assert res == [None] * world, res
    for g in members:
        for name in SHAPES:
            assert [_resolve(C, g, name, r) for r in range(world)] == [x.buffers[name].local.data_ptr() for x in members]
    for g in members:
        if g.rank == dead:
            continue
        for name in SHAPES:
            for r in range(world):
                same = g.buffers[name].peer_ptrs[r] == before[(g.rank, name)][r]
                assert same == (r != dead)             # survivor<->survivor mappings untouched, rank 2 re-opened
                if r == dead:
                    assert before[(g.rank, name)][r] in C.closed
    assert not new.deferred


def test_mismatched_newcomer_is_refused_by_all_ranks_before_any_mapping_changes(fake):
    C, world, dead = fake, 3, 1
    groups = _group_of(world, C, SHAPES)
    for name in SHAPES:
        C.allocs[groups[dead].buffers[name].local.data_ptr()] = False
    new = pm.PeerGroup.joining(dead, world, dead)
    new.alloc("model.params", 1 << 19, torch.bfloat16)            # a different model
    new.alloc("outer.flags", 64, torch.float32)
    members = [groups[0], new, groups[2]]
    before = [list(g.buffers["model.params"].peer_ptrs) for g in members]
    make = _barrier_exchange(world)
    res = _in_threads([lambda g=g: g.rebind(make(g.rank), [dead]) for g in members])
    assert all(isinstance(e, pm.RebindError) and "model.params" in str(e) for e in res), res
    assert [list(g.buffers["model.params"].peer_ptrs) for g in members] == before and not C.closed


def test_nvls_buffers_reach_a_newcomer_as_file_descriptors(fake, monkeypatch):
    """Survivors whose buffers are NVLS (VMM) allocations have no CUDA-IPC handle to offer: a replacement process
    receives one POSIX file descriptor per buffer and survivor over AF_UNIX (real sockets, pipe fds standing in
    for `cuMemExportToShareableHandle`), maps them, and the group marks its multicast team as broken so the outer
    step runs the P2P variant from then on.  The newcomer's own buffers are plain IPC and are opened as before."""
    import os

    from hypha_b200.parallel import nvls

    C, world, dead = fake, 3, 1
    groups = _group_of(world, C, SHAPES)
    pipes = {}
    for g in groups:
        for name, b in g.buffers.items():
            b.mc_ptr, b._h_local, b.nbytes_mapped = 0xabc000 + g.rank, object(), b.numel * 2

            def export_fd(g=g, name=name):
                r, w = os.pipe()
                os.write(w, f"{g.rank}:{name}".encode())
                os.close(w)
                pipes[(g.rank, name)] = r
                return r
            b.export_fd = export_fd
    for name in SHAPES:
        C.allocs[groups[dead].buffers[name].local.data_ptr()] = False
    new = pm.PeerGroup.joining(dead, world, dead)
    for name, (n, dt) in SHAPES.items():
        new.alloc(name, n, dt)
    mapped = []

    def fake_map(fd, nbytes, device):
        what = os.read(fd, 64).decode()
        mapped.append((what, nbytes, device))
        return 0x7000000 + len(mapped), ("va", nbytes, what)
    monkeypatch.setattr(nvls, "map_peer_fd", fake_map)
    monkeypatch.setattr(nvls, "unmap_peer", lambda m: None)
    members = [groups[0], new, groups[2]]
    make = _barrier_exchange(world)
    res = _in_threads([lambda g=g: g.rebind(make(g.rank), [dead]) for g in members])
    assert res == [None] * 3, res
    assert sorted(m[0] for m in mapped) == sorted(f"{r}:{n}" for r in (0, 2) for n in SHAPES)
    assert all(m[2] == dead for m in mapped)
    assert all(g.multicast_broken for g in members)
    for r in (0, 2):                                   # survivors opened the newcomer's plain IPC buffers
        for name in SHAPES:
            assert _resolve(C, groups[r], name, dead) == new.buffers[name].local.data_ptr()
    assert len(new._vmm_peer_maps) == 4 and not new.deferred

def test_store_exchange_over_a_tcp_store():
    """All-gather rounds over a TCPStore: no process group, usable by a later membership generation."""
    import socket

    import torch.distributed as dist

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    world = 3
    master = dist.TCPStore("127.0.0.1", port, world, True, wait_for_workers=False)

    def rank_fn(r):
        store = master if r == 0 else dist.TCPStore("127.0.0.1", port, world, False)
        ex = pm.store_exchange(store, r, world, prefix="gen7")
        a = ex({"rank": r, "blob": b"x" * (r + 1)})
        b = ex(r * 10)
        return a, b
    res = _in_threads([lambda r=r: rank_fn(r) for r in range(world)])
    for a, b in res:
        assert [x["rank"] for x in a] == [0, 1, 2] and [len(x["blob"]) for x in a] == [1, 2, 3]
        assert b == [0, 10, 20]


def test_rebind_with_a_vacant_slot(fake):
    """4 slots, rank 1 and rank 3 died, only rank 3 is replaced: rank 1's slot stays vacant."""
    C, world = fake, 4
    groups = _group_of(world, C, SHAPES)
    for d in (1, 3):
        for name in SHAPES:
            C.allocs[groups[d].buffers[name].local.data_ptr()] = False
    new = pm.PeerGroup.joining(3, world, 3)
    for name, (n, dt) in SHAPES.items():
        new.alloc(name, n, dt)
    members = [groups[0], groups[2], new]
    slots = [0, 2, 3]
    make = _barrier_exchange(3)
    res = _in_threads([lambda g=g, i=i: g.rebind(make(i), [3], members=slots) for i, g in enumerate(members)])
    assert res == [None] * 3, res
    for g in members:
        for name in SHAPES:
            for x in members:
                assert _resolve(C, g, name, x.rank) == x.buffers[name].local.data_ptr()
    assert new.buffers["model.params"].peer_ptrs[1] == 0          # nothing mapped for the vacant slot


def test_trainer_engine_admits_a_replacement_process(fake, monkeypatch):
    """`_TrainerEngine.regroup` + `_admit` (HYPHA_GPU_JOIN=1) with the real DiLoCoOuter membership methods, a
    real TCPStore rendezvous and the fake IPC layer: 4 slots, rank 2's process is replaced.  Survivors hand the
    shard back and keep their mutual mappings; the newcomer maps everyone, takes the survivors' round and flag
    epoch and reseeds from the lowest surviving rank."""
    import socket

    from hypha_b200.executor import training as T
    from hypha_b200.parallel.diloco import DiLoCoOuter

    monkeypatch.setenv("HYPHA_GPU_JOIN", "1")
    monkeypatch.setattr(T.torch.cuda, "synchronize", lambda *a, **k: None)
    C, world, dead = fake, 4, 2
    groups = _group_of(world, C, SHAPES)
    peers = ["w0", "w1", "w2", "w3"]
    reseeded, handed = [], []

    def make_outer(group, rank, round_no, attempt, members):
        o = object.__new__(DiLoCoOuter)
        o.world, o.rank, o.group, o.round, o.attempt, o.backend = world, rank, group, round_no, attempt, "fused"
        o.n, o.segments = 4096 * world, {rank: ("own",)}
        o.set_members(members)
        o.reseed_from = lambda src, r=rank: reseeded.append((r, src))
        # the state hand-back itself moves device memory (tests/mgpu_fault_check.py runs it on GPUs); here only
        # that every rank of the new generation enters it for the newcomer, after the owner map is known
        o.handback = lambda ex, ranks, r=rank: handed.append((r, list(ranks), dict(o.owners()).get(dead)))
        return o

    def make_engine(outer, joining, me):
        e = object.__new__(T._TrainerEngine)
        e.trainer = types.SimpleNamespace(outer=outer)
        e.joining, e._pending, e._mem, e._me = joining, [], None, me
        return e

    # survivors have already dropped rank 2 (its lease was lost a round ago); rank 3 adopted its shard
    engines = {}
    for r in (0, 1, 3):
        o = make_outer(groups[r], r, round_no=5, attempt=17 + r, members=[0, 1, 3])
        o._owner = {0: 0, 1: 1, 2: 3, 3: 3}          # what `drop_ranks([2])` left behind on every survivor
        if r == 3:
            o.segments[dead] = ("adopted",)
        engines[r] = make_engine(o, False, peers[r])
    for name in SHAPES:
        C.allocs[groups[dead].buffers[name].local.data_ptr()] = False
    new_group = pm.PeerGroup.joining(dead, world, dead)
    for name, (n, dt) in SHAPES.items():
        new_group.alloc(name, n, dt)
    engines[dead] = make_engine(make_outer(new_group, dead, round_no=0, attempt=0, members=[0, 1, 2, 3]), True, "w9")

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mem = {"generation": 3, "members": ["w0", "w1", "w3", "w9"], "joined": ["w9"], "round": 5,
           "rendezvous": f"tcp://127.0.0.1:{port}", "slots": {"w0": 0, "w1": 1, "w3": 3, "w9": 2}}
    coll = {"peers": peers, "peer": None}
    stepped = []

    def run(r):
        e = engines[r]
        e.regroup(mem, coll)
        if not e.joining:
            assert e._pending == [dead]
            e.trainer.outer.step_resilient = lambda r=r: stepped.append(r)      # the survivors' own outer step
        e.outer_step(6)
        return True
    res = _in_threads([lambda r=r: run(r) for r in (0, 1, 2, 3)])
    assert res == [True] * 4, res
    assert sorted(stepped) == [0, 1, 3] and reseeded == [(dead, 0)]
    assert sorted(handed) == [(r, [dead], 3) for r in range(4)], handed     # everyone, holder = rank 3
    members = {0: groups[0], 1: groups[1], 2: new_group, 3: groups[3]}
    for r, e in engines.items():
        o = e.trainer.outer
        assert o.members() == [0, 1, 2, 3] and o.round == 5 and o.attempt == 20 and not e.joining and not e._pending
        for name in SHAPES:
            assert [_resolve(C, members[r], name, x) for x in range(world)] == \
                   [members[x].buffers[name].local.data_ptr() for x in range(world)]
    assert dead not in engines[3].trainer.outer.segments                       # the adopter gave the shard back


def test_outer_membership_bookkeeping_on_cpu():
    """`DiLoCoOuter` drop / adopt / readmit without a GPU: every shard has exactly one owner at all times, the
    adopter is the next live rank in cyclic order, adopted shards are re-derived from the common anchor, and a
    re-admitted rank gets its shard back."""
    from hypha_b200.parallel.diloco import DiLoCoOuter, shard_bounds

    world, n = 5, 5 * 4096
    anchor = torch.arange(n, dtype=torch.float32).to(torch.bfloat16)

    def outer_of(rank):
        o = object.__new__(DiLoCoOuter)
        o.world, o.rank, o.n, o.backend = world, rank, n, "fused"
        o.anchor_buf = types.SimpleNamespace(local=anchor)
        b, l = shard_bounds(n, world, rank)
        o.segments = {rank: (b, l, anchor[b:b + l].float(), torch.zeros(l))}
        o.set_members(list(range(world)))
        return o

    outers = {r: outer_of(r) for r in range(world)}

    def owners():
        own = {}
        for r, o in outers.items():
            if r in o.members():                      # a dropped rank's own view does not count
                for k in o.segments:
                    own.setdefault(k, []).append(r)
        return own

    assert owners() == {r: [r] for r in range(world)}
    live = [0, 1, 2, 3, 4]

    def exactly_one_live_owner():
        own = owners()
        assert sorted(own) == list(range(world)), f"shards without an owner: {sorted(set(range(world)) - set(own))}"
        assert all(len(v) == 1 and v[0] in live for v in own.values()), own
        maps = [outers[r].owners() for r in live]
        assert all(m == maps[0] for m in maps)                         # replicated bookkeeping agrees
        assert {k: v[0] for k, v in own.items()} == maps[0]            # ... and matches who holds the segments

    # cascaded failures: 3 dies (4 adopts its shard), then 4 dies (0 must take over BOTH 4's shard and the
    # adopted shard 3), then 0 dies (1 inherits 0, 3 and 4)
    for dead in ([3], [4], [0]):
        live = [r for r in live if r not in dead]
        for r in live:
            outers[r].drop_ranks(dead)
        for r in dead:
            outers[r].set_members([])                 # that process is gone
        exactly_one_live_owner()
    assert sorted(outers[1].segments) == [0, 1, 3, 4]
    # rank 4 comes back (fresh state: only its own shard), takes shard 4 back; 3 stays where it is
    outers[4] = outer_of(4)
    outers[4]._owner = dict(outers[1].owners())       # what `admit` hands a newcomer
    live = [1, 2, 4]
    for r in live:
        outers[r].set_members([1, 2] if r != 4 else live)
        outers[r].readmit(4, handback=False)        # bookkeeping only: no device, no control plane here
    exactly_one_live_owner()
    assert sorted(outers[1].segments) == [0, 1, 3] and sorted(outers[4].segments) == [4]
    # a later failure must not hand shard 3 to a second owner (its cyclic successor 4 is alive again)
    live = [1, 4]
    for r in live:
        outers[r].drop_ranks([2])
    outers[2].set_members([])
    exactly_one_live_owner()
    assert outers[1].owners()[3] == 1 and outers[4].owners()[2] == 4


def test_outer_step_launch_shape_rules(monkeypatch):
    """Host-side launch decisions of the fused outer step (tuned on 8xB200, profiles/outer_sweep_n8_v2.json): one
    chunk per 128 MB of bf16 delta (at most 16) from 4 ranks up, no pipelining at N<=2; 32 link CTAs with NVLS, 48
    with P2P, 66 at N<=2; multicast only when both buffers are NVLS-bound AND every rank takes part."""
    from hypha_b200.parallel.diloco import DiLoCoConfig, DiLoCoOuter

    monkeypatch.delenv("HYPHA_OUTER_CHUNKS", raising=False)
    monkeypatch.delenv("HYPHA_OUTER_LINK_CTAS", raising=False)

    def outer(world, n, mc=False, members=None):
        o = object.__new__(DiLoCoOuter)
        o.world, o.rank, o.n, o.cfg, o.use_nvls = world, 0, n, DiLoCoConfig(), world >= 4
        o.delta_buf = types.SimpleNamespace(mc_ptr=0xd000 if mc else 0)
        o.param_buf = types.SimpleNamespace(mc_ptr=0xe000 if mc else 0)
        o.set_members(members if members is not None else list(range(world)))
        return o

    llama8b = 8_030_261_248
    assert outer(2, llama8b)._pipeline_shape() == (1, 66)
    assert outer(8, llama8b, mc=True)._pipeline_shape() == (16, 32)
    assert outer(8, llama8b, mc=False)._pipeline_shape() == (16, 48)
    assert outer(4, 64 << 20)._pipeline_shape() == (1, 48)                      # 128 MB of delta: one chunk
    assert outer(4, 320 << 20)._pipeline_shape()[0] == 5
    assert outer(8, llama8b, mc=True)._multicast_ptrs() == (0xd000, 0xe000)
    assert outer(8, llama8b, mc=True, members=[0, 1, 2, 3, 4, 5, 6])._multicast_ptrs() == (0, 0)   # a rank is missing
    assert outer(8, llama8b, mc=True, members=[0, 1, 2, 3, 4, 5, 6])._pipeline_shape() == (16, 48)
    assert outer(2, llama8b, mc=True)._multicast_ptrs() == (0, 0)               # NVLS is not used below 4 ranks
    monkeypatch.setenv("HYPHA_OUTER_CHUNKS", "3")
    monkeypatch.setenv("HYPHA_OUTER_LINK_CTAS", "20")
    assert outer(8, llama8b)._pipeline_shape() == (3, 20)


def test_fused_outer_state_includes_the_anchor():
    """Resuming the fused engine mid-round in a NEW process: the anchor (weights after the last outer step) is
    part of the outer state -- without it the next pseudo-gradient would be taken against the fresh process's
    initial weights."""
    from hypha_b200.parallel.diloco import DiLoCoOuter

    n = 4096

    def outer(params, anchor, backend="fused"):
        o = object.__new__(DiLoCoOuter)
        o.backend, o.round, o.shard_begin, o.world, o.rank = backend, 7, 0, 1, 0
        o.theta_outer, o.momentum = torch.randn(n), torch.randn(n)
        o.opt = types.SimpleNamespace(params=params)
        o.anchor_buf = types.SimpleNamespace(local=anchor)
        return o

    trained = torch.randn(n).to(torch.bfloat16)
    a = outer(trained + 1, trained.clone())                               # mid-round: params moved away from the anchor
    sd = {k: v.clone() for k, v in a.state_dict().items()}
    assert torch.equal(sd["anchor"], trained)
    b = outer(torch.zeros(n, dtype=torch.bfloat16), torch.zeros(n, dtype=torch.bfloat16))   # a fresh process
    b.load_state_dict(sd)
    assert torch.equal(b.anchor_buf.local, trained) and b.round == 7 and torch.equal(b.theta_outer, a.theta_outer)
    # old checkpoints (no anchor): treated as taken at a round boundary
    del sd["anchor"]
    c = outer(trained.clone(), torch.zeros(n, dtype=torch.bfloat16))
    c.load_state_dict(sd)
    assert torch.equal(c.anchor_buf.local, trained)
    assert "anchor" not in outer(trained, trained, backend="local").state_dict()
