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


def test_nvls_buffers_cannot_be_rebound(fake):
    g = pm.PeerGroup(0, 2, 0, lambda o: [o, o])
    g.buffers["model.params"] = types.SimpleNamespace(mc_ptr=0xdead, numel=4, dtype=torch.bfloat16)
    with pytest.raises(pm.RebindError, match="multicast"):
        g.rebind(lambda o: [o, o], [1])


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
