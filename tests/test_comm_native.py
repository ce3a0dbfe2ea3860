"""Native comm substrate (`csrc/comm/vmm.cpp`): membership-epoch table under concurrent readers, descriptor passing
over AF_UNIX; on a GPU the real VMM export / import / map round trip and a one-device multicast team."""
import os
import shutil
import tempfile
import threading

import pytest
import torch

from hypha_b200.ops.native import load_C


def test_epoch_table_drop_admit_ownership():
    C = load_C()
    t = C.EpochTable(4)
    s0 = t.snapshot()
    assert s0 == {"epoch": 0, "flag_epoch": 0, "mask": 0b1111, "world": 4, "owner": [0, 1, 2, 3]}
    assert t.drop(2, 3) == 1
    assert t.snapshot()["owner"] == [0, 1, 3, 3] and t.snapshot()["mask"] == 0b1011
    # the adopter dies later: everything it held (its own shard AND the adopted one) moves on
    assert t.drop(3, 0) == 2
    assert t.snapshot()["owner"] == [0, 1, 0, 0] and t.snapshot()["mask"] == 0b0011
    assert t.admit(2, True) == 3
    assert t.snapshot()["owner"] == [0, 1, 2, 0]
    t.admit(3, False)                                # joins as a member, shard handed back later
    assert t.snapshot()["owner"] == [0, 1, 2, 0] and t.snapshot()["mask"] == 0b1111
    t.set_flag_epoch(77)
    assert t.snapshot()["flag_epoch"] == 77 and t.snapshot()["epoch"] == 5
    with pytest.raises(RuntimeError):
        t.drop(1, 1)
    with pytest.raises(RuntimeError):
        t.admit(1)
    with pytest.raises(RuntimeError):
        C.EpochTable(9)


@pytest.mark.parametrize("world,readers", [(2, 2), (8, 6)])
def test_epoch_table_rebind_under_concurrent_launches(world, readers):
    """C++ stress: reader threads take snapshots back to back (what every kernel launch does) while a writer drops and
    re-admits ranks; no snapshot may show a shard owned by a non-member, a torn mask / epoch pair or time running back."""
    assert load_C().epoch_stress(world, readers, 50_000) == 0


def test_native_descriptor_passing_roundtrip_and_timeouts():
    C = load_C()
    d = tempfile.mkdtemp(prefix="hbc", dir="/tmp")
    try:
        path = os.path.join(d, "r0.sock")
        srv = C.uds_listen(path, 4)
        box = {}
        t = threading.Thread(target=lambda: box.update(C.uds_recv_fds(srv, 2, 3, 10.0)))
        t.start()
        pipes = {1: os.pipe(), 5: os.pipe()}
        C.uds_send_fds(path, 1, [pipes[1][1], pipes[1][0]], 5.0, 5.0)
        C.uds_send_fds(path, 5, [pipes[5][1]], 5.0, 5.0)
        t.join()
        assert sorted(box) == [1, 5] and len(box[1]) == 2 and len(box[5]) == 1
        os.write(box[5][0], b"ping")
        assert os.read(pipes[5][0], 4) == b"ping"
        for fds in box.values():
            for fd in fds:
                os.close(fd)
        for r, w in pipes.values():
            os.close(r)
            os.close(w)
        # nobody sends: the receiver gives up after its timeout instead of hanging the rank
        with pytest.raises(RuntimeError, match="timed out"):
            C.uds_recv_fds(srv, 1, 1, 0.2)
        C.uds_close(srv, path)
        assert not os.path.exists(path)
        # nobody listens: the sender gives up as well
        with pytest.raises(RuntimeError, match="never appeared"):
            C.uds_send_fds(os.path.join(d, "absent.sock"), 0, [0], 0.2, 0.2)
    finally:
        shutil.rmtree(d, ignore_errors=True)


def test_vmm_reports_missing_driver_cleanly():
    C = load_C()
    if C.vmm_driver_available():
        pytest.skip("a CUDA driver is present")
    assert C.vmm_multicast_supported(0) is False and C.vmm_usable(0, 2) is False
    with pytest.raises(RuntimeError):
        C.vmm_create(0, 1 << 20, 2)


@pytest.mark.gpu
def test_vmm_export_import_map_roundtrip_on_device():
    """cuMemCreate -> POSIX fd -> import -> second mapping of the same physical memory: a write through one virtual
    address is read back through the other (what a peer process does with a survivor's buffer)."""
    from hypha_b200.parallel.peer_memory import tensor_from_ptr

    C = load_C()
    torch.cuda.set_device(0)
    torch.zeros(1, device="cuda")
    assert C.vmm_driver_available()
    before = C.vmm_live_handles()
    hid, size, gran = C.vmm_create(0, 3 << 20, 0)
    assert size >= 3 << 20 and size % gran == 0
    p0 = C.vmm_map(hid, size, gran, 0)
    fd = C.vmm_export_fd(hid)
    hid2 = C.vmm_import_fd(fd)
    os.close(fd)
    p1 = C.vmm_map(hid2, size, gran, 0)
    assert p0 != p1
    n = size // 4
    a, b = tensor_from_ptr(p0, n, torch.float32, 0), tensor_from_ptr(p1, n, torch.float32, 0)
    a.copy_(torch.arange(n, dtype=torch.float32, device="cuda"))
    torch.cuda.synchronize()
    assert torch.equal(b.cpu(), torch.arange(n, dtype=torch.float32))
    del a, b
    C.vmm_release(hid2)
    C.vmm_release(hid)
    C.vmm_release(hid)                               # idempotent
    assert C.vmm_live_handles() == before


@pytest.mark.gpu
def test_vmm_single_device_multicast_team():
    """A multicast team of one device: bind, map the multicast view, store through it, read the unicast view."""
    from hypha_b200.parallel.peer_memory import tensor_from_ptr

    C = load_C()
    torch.cuda.set_device(0)
    torch.zeros(1, device="cuda")
    if not C.vmm_usable(0, 1):
        pytest.skip("NVSwitch multicast is not available to this process")
    hid, size, gran = C.vmm_create(0, 1 << 20, 1)
    mc = C.vmm_mc_create(1, size)
    C.vmm_mc_add_device(mc, 0)
    C.vmm_mc_bind(mc, hid, size)
    uni = C.vmm_map(hid, size, gran, 0)
    mcp = C.vmm_map(mc, size, gran, 0)
    u = tensor_from_ptr(uni, size // 4, torch.float32, 0)
    u.zero_()
    torch.cuda.synchronize()
    m = tensor_from_ptr(mcp, size // 4, torch.float32, 0)
    m.fill_(3.0)                                     # plain stores to a multicast address reach every bound copy
    torch.cuda.synchronize()
    assert float(u[0]) == 3.0 and float(u[-1]) == 3.0
    del u, m
    for h in (mc, hid):
        C.vmm_release(h)
