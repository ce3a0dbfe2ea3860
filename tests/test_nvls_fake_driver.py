"""`parallel/nvls.py: alloc_nvls` against a fake VMM layer: N "ranks" are threads of this process, file
descriptors are real (memfd) and really travel over AF_UNIX with SCM_RIGHTS through the NATIVE descriptor-passing code
(`csrc/comm/vmm.cpp`), the control-plane exchange is a
thread barrier.  Checks the collective choreography (who creates / exports / imports / binds what, in which
order relative to the exchanges) and that an injected driver failure on one rank ends as NvlsUnavailable on
EVERY rank with the exchanges still aligned -- the trainer relies on that to fall back to CUDA-IPC buffers
(the real driver calls are exercised on GPUs by tests/test_multigpu.py)."""
import itertools
import os
import threading
import types

import pytest
import torch

from hypha_b200.parallel import nvls


class FakeC:
    """Stand-in for the native VMM layer (`csrc/comm/vmm.cpp`): allocations are memfds, so export / import really go
    through descriptors; the AF_UNIX descriptor passing is the REAL native code.  Every call is logged per device."""

    def __init__(self, fail=None):
        from hypha_b200.ops.native import load_C

        self.real = load_C()
        self.fail = fail or {}                       # {(call name, device): True}
        self.lock = threading.Lock()
        self.handles = {}                            # registry id -> {"kind", "obj", ...}
        self.objects = {}                            # object id (shared between importers) -> state
        self.ids = itertools.count(100)
        self.va = itertools.count(1 << 30, 1 << 24)
        self.log = []
        self.tls = threading.local()
        self.uds_listen = self.real.uds_listen
        self.uds_close = self.real.uds_close
        self.uds_send_fds = self.real.uds_send_fds
        self.uds_recv_fds = self.real.uds_recv_fds

    def _dev(self):
        return getattr(self.tls, "dev", -1)

    def _call(self, name):
        with self.lock:
            self.log.append((self._dev(), name))
        if self.fail.get((name, self._dev())):
            raise self.real.VmmUnavailable(f"{name} failed: CUDA_ERROR_NOT_PERMITTED (800)")

    def vmm_multicast_supported(self, d):
        self.tls.dev = d
        return True

    def vmm_create(self, device, nbytes, world):
        self._call("vmm_create")
        gran = 2 << 20
        size = (nbytes + gran - 1) // gran * gran
        with self.lock:
            oid, hid = next(self.ids), next(self.ids)
            self.objects[oid] = {"kind": "mem", "fd": os.memfd_create(f"mem{oid}"), "dev": device}
            self.handles[hid] = {"obj": oid, "mapped": False}
        return hid, size, gran

    def vmm_mc_create(self, world, size):
        self._call("vmm_mc_create")
        with self.lock:
            oid, hid = next(self.ids), next(self.ids)
            self.objects[oid] = {"kind": "mc", "fd": os.memfd_create(f"mc{oid}"), "devs": set(), "bound": set(), "n": world}
            self.handles[hid] = {"obj": oid, "mapped": False}
        return hid

    def vmm_export_fd(self, hid):
        self._call("vmm_export_fd")
        oid = self.handles[hid]["obj"]
        fd = os.dup(self.objects[oid]["fd"])
        os.lseek(fd, 0, os.SEEK_SET)
        os.write(fd, str(oid).encode().ljust(16))    # so that the importing side can find the object again
        return fd

    def vmm_import_fd(self, fd, multicast=False):
        self._call("vmm_import_fd")
        os.lseek(fd, 0, os.SEEK_SET)
        oid = int(os.read(fd, 16).split()[0])
        with self.lock:
            hid = next(self.ids)
            self.handles[hid] = {"obj": oid, "mapped": False}
        return hid

    def vmm_map(self, hid, size, gran, device):
        self._call("vmm_map")
        self.handles[hid]["mapped"] = True
        return next(self.va)

    def vmm_release(self, hid):
        with self.lock:
            h = self.handles.pop(hid, None)
            self.log.append((self._dev(), "vmm_release"))
            if h is not None and h["mapped"]:
                self.log.append((self._dev(), "unmapped"))

    def vmm_mc_add_device(self, mc, dev):
        self._call("vmm_mc_add_device")
        with self.lock:
            self.objects[self.handles[mc]["obj"]]["devs"].add(dev)

    def vmm_mc_bind(self, mc, mem, size):
        self._call("vmm_mc_bind")
        with self.lock:
            o = self.objects[self.handles[mc]["obj"]]
            assert len(o["devs"]) == o["n"], "bind before every device joined the team"
            o["bound"].add(self.objects[self.handles[mem]["obj"]]["dev"])


class FakeGroup:
    """PeerGroup stand-in: `_exchange` is an all-gather over a thread barrier."""

    def __init__(self, rank, world, shared):
        self.rank, self.world, self.device, self.shared = rank, world, rank, shared
        self.buffers = {}
        self.exchanges = 0

    def _exchange(self, obj):
        s = self.shared
        self.exchanges += 1
        s["slots"][self.rank] = obj
        s["barrier"].wait(timeout=30)
        out = list(s["slots"])
        s["barrier"].wait(timeout=30)
        return out


@pytest.fixture
def sock_dir():
    """AF_UNIX paths are limited to 107 bytes: keep the sockets out of pytest's (possibly long) basetemp."""
    import shutil
    import tempfile

    d = tempfile.mkdtemp(prefix="hbn", dir="/tmp")
    yield d
    shutil.rmtree(d, ignore_errors=True)


def _run(world, fake, tmp, monkeypatch):
    monkeypatch.setattr(nvls, "load_C", lambda: fake)
    monkeypatch.setattr(nvls, "tensor_from_ptr", lambda ptr, numel, dtype, device: torch.zeros(numel, dtype=dtype))
    fake_cuda = types.SimpleNamespace(set_device=lambda d: None, synchronize=lambda: None,
                                      current_stream=lambda: types.SimpleNamespace(synchronize=lambda: None))
    monkeypatch.setattr(nvls, "torch", types.SimpleNamespace(cuda=fake_cuda, dtype=torch.dtype, bfloat16=torch.bfloat16))
    shared = {"slots": [None] * world, "barrier": threading.Barrier(world)}
    groups = [FakeGroup(r, world, shared) for r in range(world)]
    results = [None] * world

    def worker(r):
        try:
            results[r] = nvls.alloc_nvls(groups[r], "model.params", 1 << 20, torch.bfloat16, sock_dir=tmp)
        except Exception as exc:      # noqa: BLE001 - the test inspects the type
            results[r] = exc

    threads = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(120)
    assert not any(t.is_alive() for t in threads), "a rank hung"
    return groups, results


@pytest.mark.parametrize("world", [2, 4])
def test_alloc_nvls_choreography(world, sock_dir, monkeypatch):
    fake = FakeC()
    groups, results = _run(world, fake, sock_dir, monkeypatch)
    assert all(isinstance(b, nvls.NvlsBuffer) for b in results), results
    for r, b in enumerate(results):
        assert b.mc_ptr and len(b.peer_ptrs) == world and len(set(b.peer_ptrs)) == world and b.local.numel() == 1 << 20
        assert groups[r].buffers["model.params"] is b
    mcs = [o for o in fake.objects.values() if o["kind"] == "mc"]
    assert len(mcs) == 1 and mcs[0]["devs"] == set(range(world)) and mcs[0]["bound"] == set(range(world))
    assert sum(1 for d, n in fake.log if n == "vmm_mc_create") == 1 and fake.log.count((0, "vmm_mc_create")) == 1
    assert sum(1 for d, n in fake.log if n == "vmm_create") == world
    per_rank_imports = [sum(1 for d, n in fake.log if d == r and n == "vmm_import_fd") for r in range(world)]
    assert per_rank_imports == [world - 1] + [world] * (world - 1)      # peers' memory (+ the multicast object)
    assert len({g.exchanges for g in groups}) == 1                       # every rank made the same number of exchanges
    assert not [f for f in os.listdir(sock_dir) if f.endswith(".sock")]
    for b in results:
        b.close()
        assert b.mc_ptr == 0
    assert not fake.handles                                             # every registry entry was released


@pytest.mark.parametrize("call,dev,exchanges", [
    ("vmm_mc_create", 0, 1),                  # rank 0 cannot create the team object      -> agreed at the first exchange
    ("vmm_create", 2, 1),                     # one rank is out of memory                    -> first exchange
    ("vmm_import_fd", 1, 3),                  # a peer's handle does not import            -> after the swap
    ("vmm_mc_add_device", 2, 3),
    ("vmm_mc_bind", 1, 4),                    # binding fails on one rank                    -> nobody maps the view
])
def test_alloc_nvls_failure_is_agreed_by_all_ranks(call, dev, exchanges, sock_dir, monkeypatch):
    world = 3
    fake = FakeC(fail={(call, dev): True})
    groups, results = _run(world, fake, sock_dir, monkeypatch)
    assert all(isinstance(e, nvls.NvlsUnavailable) for e in results), results
    assert all(f"rank {dev}" in str(e) and call in str(e) for e in results), results
    assert [g.exchanges for g in groups] == [exchanges] * world         # still aligned: the group stays usable
    assert not any(g.buffers for g in groups)
    assert not [f for f in os.listdir(sock_dir) if f.endswith(".sock")]
    # what was mapped before the failure has been unmapped again, and no registry entry is left behind
    maps = sum(1 for d, n in fake.log if n == "vmm_map")
    unmaps = sum(1 for d, n in fake.log if n == "unmapped")
    assert maps == unmaps and not fake.handles
