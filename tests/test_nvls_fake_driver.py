"""`parallel/nvls.py: alloc_nvls` against a fake CUDA driver: N "ranks" are threads of this process, file
descriptors are real (memfd) and really travel over AF_UNIX with SCM_RIGHTS, the control-plane exchange is a
thread barrier.  Checks the collective choreography (who creates / exports / imports / binds what, in which
order relative to the exchanges) and that an injected driver failure on one rank ends as NvlsUnavailable on
EVERY rank with the exchanges still aligned -- the trainer relies on that to fall back to CUDA-IPC buffers
(the real driver calls are exercised on GPUs by tests/test_multigpu.py)."""
import itertools
import os
import threading
import types
from enum import Enum

import pytest
import torch

from hypha_b200.parallel import nvls


class _R(Enum):
    CUDA_SUCCESS = 0
    CUDA_ERROR_NOT_PERMITTED = 800


class _E(Enum):
    A = 1


class FakeCu:
    """Just enough of cuda.bindings.driver; every call is logged per device."""

    CUresult = _R
    CUmemAllocationHandleType = types.SimpleNamespace(CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR=1)
    CUmemAllocationType = types.SimpleNamespace(CU_MEM_ALLOCATION_TYPE_PINNED=1)
    CUmemLocationType = types.SimpleNamespace(CU_MEM_LOCATION_TYPE_DEVICE=1)
    CUmulticastGranularity_flags = types.SimpleNamespace(CU_MULTICAST_GRANULARITY_RECOMMENDED=1)
    CUmemAllocationGranularity_flags = types.SimpleNamespace(CU_MEM_ALLOC_GRANULARITY_RECOMMENDED=1)
    CUmemAccess_flags = types.SimpleNamespace(CU_MEM_ACCESS_FLAGS_PROT_READWRITE=3)
    CUdevice_attribute = types.SimpleNamespace(CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED=1)

    def __init__(self, fail=None):
        self.fail = fail or {}                       # {(call name, device): True}
        self.lock = threading.Lock()
        self.handles = {}                            # handle id -> {"kind", "fd", "dev"}
        self.ids = itertools.count(100)
        self.va = itertools.count(1 << 30, 1 << 24)
        self.log = []
        self.tls = threading.local()

    def _dev(self):
        return getattr(self.tls, "dev", -1)

    def _ret(self, name, *vals):
        with self.lock:
            self.log.append((self._dev(), name))
        if self.fail.get((name, self._dev())):
            return (_R.CUDA_ERROR_NOT_PERMITTED,) + tuple(None for _ in vals)
        return (_R.CUDA_SUCCESS,) + vals

    @staticmethod
    def _ns():
        return types.SimpleNamespace(location=types.SimpleNamespace())

    CUmemAllocationProp = staticmethod(lambda: FakeCu._ns())
    CUmulticastObjectProp = staticmethod(lambda: types.SimpleNamespace())
    CUmemAccessDesc = staticmethod(lambda: FakeCu._ns())

    def cuInit(self, flags): return self._ret("cuInit")
    def cuDeviceGet(self, d):
        self.tls.dev = d
        return self._ret("cuDeviceGet", d)
    def cuDeviceGetAttribute(self, a, d): return self._ret("cuDeviceGetAttribute", 1)
    def cuMulticastGetGranularity(self, p, f): return self._ret("cuMulticastGetGranularity", 2 << 20)
    def cuMemGetAllocationGranularity(self, p, f): return self._ret("cuMemGetAllocationGranularity", 2 << 20)

    def cuMemCreate(self, size, prop, flags):
        h = next(self.ids)
        self.handles[h] = {"kind": "mem", "fd": os.memfd_create(f"mem{h}"), "dev": prop.location.id, "size": size}
        return self._ret("cuMemCreate", h)

    def cuMulticastCreate(self, prop):
        h = next(self.ids)
        self.handles[h] = {"kind": "mc", "fd": os.memfd_create(f"mc{h}"), "devs": set(), "bound": set(), "n": prop.numDevices}
        return self._ret("cuMulticastCreate", h)

    def cuMemExportToShareableHandle(self, h, typ, flags):
        fd = os.dup(self.handles[h]["fd"])
        os.write(fd, str(h).encode().ljust(16))      # so that the importing side can find the object again
        return self._ret("cuMemExportToShareableHandle", fd)

    def cuMemImportFromShareableHandle(self, fd, typ):
        os.lseek(fd, 0, os.SEEK_SET)
        return self._ret("cuMemImportFromShareableHandle", int(os.read(fd, 16).split()[0]))

    def cuMemAddressReserve(self, size, align, addr, flags): return self._ret("cuMemAddressReserve", next(self.va))
    def cuMemMap(self, va, size, off, h, flags): return self._ret("cuMemMap")
    def cuMemSetAccess(self, va, size, desc, n): return self._ret("cuMemSetAccess")
    def cuMemUnmap(self, va, size): return self._ret("cuMemUnmap")
    def cuMemAddressFree(self, va, size): return self._ret("cuMemAddressFree")
    def cuMemRelease(self, h): return self._ret("cuMemRelease")

    def cuMulticastAddDevice(self, mc, dev):
        with self.lock:
            self.handles[mc]["devs"].add(dev)
        return self._ret("cuMulticastAddDevice")

    def cuMulticastBindMem(self, mc, mo, h, ho, size, flags):
        with self.lock:
            assert len(self.handles[mc]["devs"]) == self.handles[mc]["n"], "bind before every device joined the team"
            self.handles[mc]["bound"].add(self.handles[h]["dev"])
        return self._ret("cuMulticastBindMem")


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
    monkeypatch.setattr(nvls, "_cu", lambda: fake)
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
    fake = FakeCu()
    groups, results = _run(world, fake, sock_dir, monkeypatch)
    assert all(isinstance(b, nvls.NvlsBuffer) for b in results), results
    for r, b in enumerate(results):
        assert b.mc_ptr and len(b.peer_ptrs) == world and len(set(b.peer_ptrs)) == world and b.local.numel() == 1 << 20
        assert groups[r].buffers["model.params"] is b
    mcs = [h for h in fake.handles.values() if h["kind"] == "mc"]
    assert len(mcs) == 1 and mcs[0]["devs"] == set(range(world)) and mcs[0]["bound"] == set(range(world))
    assert sum(1 for d, n in fake.log if n == "cuMulticastCreate") == 1 and fake.log.count((0, "cuMulticastCreate")) == 1
    assert sum(1 for d, n in fake.log if n == "cuMemCreate") == world
    per_rank_imports = [sum(1 for d, n in fake.log if d == r and n == "cuMemImportFromShareableHandle") for r in range(world)]
    assert per_rank_imports == [world - 1] + [world] * (world - 1)      # peers' memory (+ the multicast object)
    assert len({g.exchanges for g in groups}) == 1                       # every rank made the same number of exchanges
    assert not [f for f in os.listdir(sock_dir) if f.endswith(".sock")]
    for b in results:
        b.close()
        assert b.mc_ptr == 0


@pytest.mark.parametrize("call,dev,exchanges", [
    ("cuMulticastCreate", 0, 1),              # rank 0 cannot create the team object      -> agreed at the first exchange
    ("cuMemCreate", 2, 1),                    # one rank is out of memory                    -> first exchange
    ("cuMemImportFromShareableHandle", 1, 3),  # a peer's handle does not import            -> after the swap
    ("cuMulticastAddDevice", 2, 3),
    ("cuMulticastBindMem", 1, 4),             # binding fails on one rank                    -> nobody maps the view
])
def test_alloc_nvls_failure_is_agreed_by_all_ranks(call, dev, exchanges, sock_dir, monkeypatch):
    world = 3
    fake = FakeCu(fail={(call, dev): True})
    groups, results = _run(world, fake, sock_dir, monkeypatch)
    assert all(isinstance(e, nvls.NvlsUnavailable) for e in results), results
    assert all(f"rank {dev}" in str(e) and call.replace("cuMemImportFromShareableHandle", "mport") in str(e) for e in results), results
    assert [g.exchanges for g in groups] == [exchanges] * world         # still aligned: the group stays usable
    assert not any(g.buffers for g in groups)
    assert not [f for f in os.listdir(sock_dir) if f.endswith(".sock")]
    # what was mapped before the failure has been unmapped again
    maps = sum(1 for d, n in fake.log if n == "cuMemMap")
    unmaps = sum(1 for d, n in fake.log if n == "cuMemUnmap")
    assert maps == unmaps
