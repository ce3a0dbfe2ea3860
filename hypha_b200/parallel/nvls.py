"""NVLS symmetric memory: VMM allocations that are (a) local HBM, (b) mapped on every peer for
plain P2P loads/stores and (c) bound to an NVSwitch *multicast object*, so a kernel can issue
``multimem.ld_reduce`` (the switch sums the N copies in flight) and ``multimem.st`` (the switch
replicates one store to all N copies).

Set-up (CUDA driver API from C++, `csrc/comm/vmm.cpp`: allocation, export / import, mapping, multicast objects and
the SCM_RIGHTS descriptor passing are native; this module only sequences the stages between control-plane exchanges;
no NCCL, no torch symmetric-memory, no cuda-python):
  1. ``cuMemCreate`` a POSIX-fd-exportable physical allocation per rank,
  2. swap the fds between the rank processes over AF_UNIX sockets (``SCM_RIGHTS``),
  3. import + ``cuMemMap`` every peer's allocation (unicast view, for flags / fallback P2P),
  4. rank 0 ``cuMulticastCreate``s the multicast object and shares its fd, every rank
     ``cuMulticastAddDevice`` + ``cuMulticastBindMem``s its allocation and maps the multicast VA.

This is the B200-native replacement of the reference's parameter-server fan-in / fan-out
(crates/worker/src/executor/parameter_server.rs:101-283) at the transport level.
"""
from __future__ import annotations

import os
import tempfile
from dataclasses import dataclass
from typing import List, Optional

import torch

from ..ops.native import load_C
from .peer_memory import _DT_SIZE, PeerGroup, SymmetricBuffer, tensor_from_ptr


class NvlsUnavailable(RuntimeError):
    pass


def _native(fn, *args):
    """Call into `csrc/comm/vmm.cpp`; its `Unavailable` (and any driver error) becomes `NvlsUnavailable`."""
    try:
        return fn(*args)
    except RuntimeError as exc:          # VmmUnavailable derives from RuntimeError
        raise NvlsUnavailable(str(exc)) from exc


def _send_fds(path: str, rank: int, fds: List[int]) -> None:
    """Hand ``fds`` to the process listening on ``path`` (SCM_RIGHTS); waits up to 30 s for the socket to appear and
    up to 60 s for the receiver's acknowledgement."""
    _native(load_C().uds_send_fds, path, int(rank), [int(f) for f in fds], 30.0, 60.0)


def _recv_fds(server, count: int, nfds: int):
    """``{sender rank: [fd, ...]}`` from ``count`` peers; ``server`` is a listening socket (object or descriptor)."""
    fd = server if isinstance(server, int) else server.fileno()
    got = _native(load_C().uds_recv_fds, fd, int(count), int(nfds), 60.0)
    return {int(r): list(f) for r, f in got.items()}


@dataclass
class NvlsBuffer(SymmetricBuffer):
    mc_ptr: int = 0                 # multicast virtual address (0 if multicast could not be set up)
    nbytes_mapped: int = 0
    _handles: Optional[list] = None  # native registry ids to release in close()
    _h_local: int = 0                # the local physical allocation (exportable as a POSIX fd: `export_fd`)
    _gran: int = 0

    def export_fd(self) -> int:
        """A POSIX file descriptor of the local allocation for a process that was not part of the original team
        (a replacement rank): it imports + maps it with `map_peer_fd` and reaches this rank's copy with plain
        P2P loads/stores.  The caller closes the descriptor after sending it."""
        return int(_native(load_C().vmm_export_fd, self._h_local))

    def close(self) -> None:
        """Unmap every view and release the allocation handles (idempotent; errors are ignored: the
        peers may already be gone at shutdown)."""
        if not self._handles:
            return
        C = load_C()
        for hid in self._handles:
            C.vmm_release(hid)               # noexcept on the native side
        self._handles = None
        self.mc_ptr = 0


def multicast_supported(device: int) -> bool:
    return bool(load_C().vmm_multicast_supported(int(device)))


def nvls_usable(device: int, world: int) -> bool:
    """Rank-local pre-flight for the NVLS path, with NO inter-rank communication: the attribute says
    multicast exists, and this process can actually create an exportable allocation and a multicast
    object for ``world`` devices and add its own device to it.  Any failure (fabric manager not running,
    multicast not permitted in this container, out of handles, ...) returns False, so that all ranks can
    agree on plain CUDA-IPC buffers *before* anyone enters the collective set-up in ``alloc_nvls``."""
    try:
        return bool(load_C().vmm_usable(int(device), max(1, int(world))))
    except Exception:
        return False


def map_peer_fd(fd: int, nbytes_mapped: int, device: int) -> tuple:
    """Import a peer's VMM allocation from its file descriptor and map it read/write on ``device``.
    Returns ``(device pointer, mapping)``; the mapping goes to `unmap_peer`."""
    C = load_C()
    hid = _native(C.vmm_import_fd, int(fd), False)
    try:
        ptr = _native(C.vmm_map, hid, int(nbytes_mapped), 0, int(device))
    except Exception:
        C.vmm_release(hid)
        raise
    return int(ptr), hid


def unmap_peer(mapping) -> None:
    load_C().vmm_release(int(mapping))


def alloc_nvls(group: PeerGroup, name: str, numel: int, dtype: torch.dtype, sock_dir: Optional[str] = None) -> NvlsBuffer:
    """Collective: every rank of ``group`` must call it with the same arguments.

    The set-up is five stages separated by control-plane exchanges.  A stage that fails on one rank does not
    raise there and then (the other ranks would block in the next exchange, or bind memory to a team with a
    missing member): the error travels WITH the next exchange, every rank sees it at the same point, undoes
    what it mapped so far and raises ``NvlsUnavailable`` -- the number of exchanges each rank performs is the
    same on every path, so the caller can go on using the group (``trainer.py`` falls back to CUDA-IPC)."""
    import threading

    C = load_C()
    rank, world, device = group.rank, group.world, group.device
    st = {"err": None, "fds": [], "ids": [], "server": -1, "path": None}

    def stage(fn) -> None:
        if st["err"] is None:
            try:
                fn()
            except Exception as exc:     # noqa: BLE001 - reported to every rank by the next exchange
                st["err"] = f"rank {rank}: {type(exc).__name__}: {exc}"

    def undo() -> None:
        for fd in st["fds"]:
            try:
                os.close(fd)
            except OSError:
                pass
        st["fds"].clear()
        if st["server"] >= 0:
            C.uds_close(st["server"], st["path"] or "")
            st["server"] = -1
        for hid in st["ids"]:
            C.vmm_release(hid)
        st["ids"].clear()

    def agree(extra=None) -> list:
        got = group._exchange((st["err"], extra))
        bad = [e for e, _ in got if e]
        if bad:
            undo()
            raise NvlsUnavailable("; ".join(bad))
        return [x for _, x in got]

    # ---- stage 1: local allocation (+ the multicast object on rank 0), exportable as POSIX fds ----------------
    def s_create() -> None:
        torch.cuda.set_device(device)
        torch.cuda.current_stream().synchronize()     # make sure the primary context exists
        if not C.vmm_multicast_supported(device):
            raise NvlsUnavailable("device does not report CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED")
        nbytes = numel * _DT_SIZE[dtype]
        st["h_local"], st["size"], st["gran"] = _native(C.vmm_create, device, nbytes, world)
        st["ids"].append(st["h_local"])
        st["fd_local"] = int(_native(C.vmm_export_fd, st["h_local"]))
        st["fds"].append(st["fd_local"])
        if rank == 0:
            st["mc"] = _native(C.vmm_mc_create, world, st["size"])
            st["ids"].append(st["mc"])
            st["fd_mc"] = int(_native(C.vmm_export_fd, st["mc"]))
            st["fds"].append(st["fd_mc"])

    stage(s_create)
    tag = agree(os.getpid() if rank == 0 else None)[0]

    # ---- stage 2: every rank listens on an AF_UNIX socket --------------------------------------------------------
    base = sock_dir or tempfile.gettempdir()
    paths = [os.path.join(base, f"hb-nvls-{tag}-{name.replace('.', '_')}-{r}.sock") for r in range(world)]

    def s_listen() -> None:
        st["path"] = paths[rank]
        st["server"] = int(_native(C.uds_listen, paths[rank], world))

    stage(s_listen)
    agree()                                        # everyone is listening

    # ---- stage 3: descriptors travel (SCM_RIGHTS), peers' memory is mapped, the device joins the team --------------
    ptrs = [0] * world

    def s_swap() -> None:
        recv_box: dict = {}
        recv_err: list = []

        def receiver() -> None:
            try:
                recv_box.update(_recv_fds(st["server"], world - 1, 2))   # a peer that dies must not hang us: 60 s
            except Exception as exc:     # noqa: BLE001
                recv_err.append(exc)

        t = threading.Thread(target=receiver)
        t.start()
        send_err = None
        for p in range(world):
            if p != rank:
                try:
                    _send_fds(paths[p], rank, [st["fd_local"]] + ([st["fd_mc"]] if rank == 0 else []))
                except Exception as exc:     # noqa: BLE001 - keep serving the other peers
                    send_err = send_err or exc
        t.join()
        C.uds_close(st["server"], paths[rank])
        st["server"] = -1
        for fds in recv_box.values():
            st["fds"].extend(fds)
        if send_err or recv_err:
            raise send_err or recv_err[0]

        size, gran = st["size"], st["gran"]
        ptrs[rank] = int(_native(C.vmm_map, st["h_local"], size, gran, device))
        for p in range(world):
            if p == rank:
                continue
            hp = _native(C.vmm_import_fd, recv_box[p][0], False)
            st["ids"].append(hp)
            ptrs[p] = int(_native(C.vmm_map, hp, size, gran, device))
        if rank != 0:
            st["mc"] = _native(C.vmm_import_fd, recv_box[0][1], True)
            st["ids"].append(st["mc"])
        for fd in st["fds"]:
            os.close(fd)
        st["fds"].clear()
        _native(C.vmm_mc_add_device, st["mc"], device)

    stage(s_swap)
    agree()                                        # every device is part of the team before anyone binds

    # ---- stage 4: bind the local memory to the multicast object -------------------------------------------------------
    stage(lambda: _native(C.vmm_mc_bind, st["mc"], st["h_local"], st["size"]))
    agree()

    # ---- stage 5: map the multicast view, zero the local slice ----------------------------------------------------------
    def s_view() -> None:
        st["mc_ptr"] = int(_native(C.vmm_map, st["mc"], st["size"], st["gran"], device))
        local = tensor_from_ptr(ptrs[rank], numel, dtype, device)
        local.zero_()
        torch.cuda.synchronize()
        st["local"] = local

    stage(s_view)
    agree()
    # release order: the multicast view and the peers' mappings first, the local allocation last
    ids = [st["mc"]] + [h for h in st["ids"] if h not in (st["mc"], st["h_local"])] + [st["h_local"]]
    buf = NvlsBuffer(name, numel, dtype, st["local"], ptrs, mc_ptr=st["mc_ptr"], nbytes_mapped=st["size"],
                     _handles=ids, _h_local=st["h_local"], _gran=st["gran"])
    group.buffers[name] = buf
    return buf
