"""NVLS symmetric memory: VMM allocations that are (a) local HBM, (b) mapped on every peer for
plain P2P loads/stores and (c) bound to an NVSwitch *multicast object*, so a kernel can issue
``multimem.ld_reduce`` (the switch sums the N copies in flight) and ``multimem.st`` (the switch
replicates one store to all N copies).

Set-up (driver API through ``cuda.bindings``; no NCCL, no torch symmetric-memory):
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
import socket
import tempfile
import time
from dataclasses import dataclass
from typing import List, Optional

import torch

from .peer_memory import _DT_SIZE, PeerGroup, SymmetricBuffer, tensor_from_ptr


class NvlsUnavailable(RuntimeError):
    pass


def _cu():
    try:
        from cuda.bindings import driver as cu
    except ImportError:  # older cuda-python
        from cuda import cuda as cu
    return cu


def _ck(res, what: str):
    cu = _cu()
    err, *rest = res if isinstance(res, tuple) else (res,)
    if err != cu.CUresult.CUDA_SUCCESS:
        raise NvlsUnavailable(f"{what} failed: {err}")
    return rest[0] if len(rest) == 1 else (tuple(rest) if rest else None)


def _send_fds(path: str, rank: int, fds: List[int]) -> None:
    deadline = time.time() + 30
    while True:
        try:
            s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            s.connect(path)
            break
        except (FileNotFoundError, ConnectionRefusedError):
            s.close()
            if time.time() > deadline:
                raise NvlsUnavailable(f"peer socket {path} never appeared")
            time.sleep(0.02)
    with s:
        s.settimeout(60)    # a receiver that failed before accepting must not hang the sender
        try:
            socket.send_fds(s, [rank.to_bytes(4, "little")], fds)
            s.recv(1)       # ack: the receiver has dup'ed the descriptors
        except (socket.timeout, OSError) as exc:
            raise NvlsUnavailable(f"peer at {path} did not take the memory handles: {exc}") from exc


def _recv_fds(server: socket.socket, count: int, nfds: int):
    got = {}
    for _ in range(count):
        try:
            conn, _ = server.accept()
        except socket.timeout as exc:
            raise NvlsUnavailable("timed out waiting for a peer's memory handle") from exc
        with conn:
            msg, fds, _, _ = socket.recv_fds(conn, 4, nfds)
            got[int.from_bytes(msg, "little")] = list(fds)
            conn.send(b"k")
    return got


@dataclass
class NvlsBuffer(SymmetricBuffer):
    mc_ptr: int = 0                 # multicast virtual address (0 if multicast could not be set up)
    nbytes_mapped: int = 0
    _handles: Optional[list] = None  # driver handles to release in close(): [(va, size, handle), ...]
    _h_local: object = None          # the local physical allocation (exportable as a POSIX fd: `export_fd`)
    _gran: int = 0

    def export_fd(self) -> int:
        """A POSIX file descriptor of the local allocation for a process that was not part of the original team
        (a replacement rank): it imports + maps it with `map_peer_fd` and reaches this rank's copy with plain
        P2P loads/stores.  The caller closes the descriptor after sending it."""
        cu = _cu()
        H = cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
        return int(_ck(cu.cuMemExportToShareableHandle(self._h_local, H, 0), "cuMemExportToShareableHandle"))

    def close(self) -> None:
        """Unmap every view and release the allocation handles (idempotent; errors are ignored: the
        peers may already be gone at shutdown)."""
        if not self._handles:
            return
        cu = _cu()
        for va, size, handle in self._handles:
            try:
                cu.cuMemUnmap(va, size)
                cu.cuMemAddressFree(va, size)
                cu.cuMemRelease(handle)
            except Exception:
                pass
        self._handles = None
        self.mc_ptr = 0


def multicast_supported(device: int) -> bool:
    cu = _cu()
    try:
        (err,) = cu.cuInit(0)
        err, dev = cu.cuDeviceGet(device)
        err, val = cu.cuDeviceGetAttribute(cu.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev)
        return err == cu.CUresult.CUDA_SUCCESS and bool(val)
    except Exception:
        return False


def nvls_usable(device: int, world: int) -> bool:
    """Rank-local pre-flight for the NVLS path, with NO inter-rank communication: the attribute says
    multicast exists, and this process can actually create an exportable allocation and a multicast
    object for ``world`` devices and add its own device to it.  Any failure (fabric manager not running,
    multicast not permitted in this container, out of handles, ...) returns False, so that all ranks can
    agree on plain CUDA-IPC buffers *before* anyone enters the collective set-up in ``alloc_nvls``."""
    if not multicast_supported(device):
        return False
    cu = _cu()
    handles = []
    fds = []
    try:
        dev = _ck(cu.cuDeviceGet(device), "cuDeviceGet")
        H = cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
        prop = cu.CUmemAllocationProp()
        prop.type = cu.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
        prop.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
        prop.location.id = device
        prop.requestedHandleTypes = H
        mcprop = cu.CUmulticastObjectProp()
        mcprop.numDevices = max(1, int(world))
        mcprop.handleTypes = H
        mcprop.flags = 0
        mcprop.size = 1
        gran = max(int(_ck(cu.cuMulticastGetGranularity(mcprop, cu.CUmulticastGranularity_flags.CU_MULTICAST_GRANULARITY_RECOMMENDED),
                           "cuMulticastGetGranularity")),
                   int(_ck(cu.cuMemGetAllocationGranularity(prop, cu.CUmemAllocationGranularity_flags.CU_MEM_ALLOC_GRANULARITY_RECOMMENDED),
                           "cuMemGetAllocationGranularity")))
        mcprop.size = gran
        h = _ck(cu.cuMemCreate(gran, prop, 0), "cuMemCreate")
        handles.append(h)
        fds.append(int(_ck(cu.cuMemExportToShareableHandle(h, H, 0), "cuMemExportToShareableHandle")))
        mc = _ck(cu.cuMulticastCreate(mcprop), "cuMulticastCreate")
        handles.append(mc)
        fds.append(int(_ck(cu.cuMemExportToShareableHandle(mc, H, 0), "export multicast handle")))
        _ck(cu.cuMulticastAddDevice(mc, dev), "cuMulticastAddDevice")
        return True
    except Exception:
        return False
    finally:
        for fd in fds:
            try:
                os.close(fd)
            except OSError:
                pass
        for h in reversed(handles):
            try:
                cu.cuMemRelease(h)
            except Exception:
                pass


def map_peer_fd(fd: int, nbytes_mapped: int, device: int) -> tuple:
    """Import a peer's VMM allocation from its file descriptor and map it read/write on ``device``.
    Returns ``(device pointer, (va, size, handle))``; the tuple goes to `unmap_peer`."""
    cu = _cu()
    H = cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
    h = _ck(cu.cuMemImportFromShareableHandle(fd, H), "cuMemImportFromShareableHandle")
    try:
        va = _ck(cu.cuMemAddressReserve(nbytes_mapped, 0, 0, 0), "cuMemAddressReserve")
        try:
            _ck(cu.cuMemMap(va, nbytes_mapped, 0, h, 0), "cuMemMap")
        except Exception:
            cu.cuMemAddressFree(va, nbytes_mapped)
            raise
        access = cu.CUmemAccessDesc()
        access.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
        access.location.id = device
        access.flags = cu.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE
        _ck(cu.cuMemSetAccess(va, nbytes_mapped, [access], 1), "cuMemSetAccess")
    except Exception:
        cu.cuMemRelease(h)
        raise
    return int(va), (va, nbytes_mapped, h)


def unmap_peer(mapping: tuple) -> None:
    cu = _cu()
    va, size, h = mapping
    try:
        cu.cuMemUnmap(va, size)
        cu.cuMemAddressFree(va, size)
        cu.cuMemRelease(h)
    except Exception:
        pass


def alloc_nvls(group: PeerGroup, name: str, numel: int, dtype: torch.dtype, sock_dir: Optional[str] = None) -> NvlsBuffer:
    """Collective: every rank of ``group`` must call it with the same arguments.

    The set-up is five stages separated by control-plane exchanges.  A stage that fails on one rank does not
    raise there and then (the other ranks would block in the next exchange, or bind memory to a team with a
    missing member): the error travels WITH the next exchange, every rank sees it at the same point, undoes
    what it mapped so far and raises ``NvlsUnavailable`` -- the number of exchanges each rank performs is the
    same on every path, so the caller can go on using the group (``trainer.py`` falls back to CUDA-IPC)."""
    import threading

    cu = _cu()
    rank, world, device = group.rank, group.world, group.device
    H = cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
    st = {"err": None, "fds": [], "mapped": [], "loose": [], "server": None, "path": None}

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
        if st["server"] is not None:
            st["server"].close()
            st["server"] = None
        if st["path"] and os.path.exists(st["path"]):
            os.unlink(st["path"])
        for va, size, h in st["mapped"]:
            try:
                cu.cuMemUnmap(va, size)
                cu.cuMemAddressFree(va, size)
            except Exception:
                pass
        for h in [m[2] for m in st["mapped"]] + st["loose"]:
            try:
                cu.cuMemRelease(h)
            except Exception:
                pass
        st["mapped"].clear()
        st["loose"].clear()

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
        if not multicast_supported(device):
            raise NvlsUnavailable("device does not report CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED")
        _ck(cu.cuInit(0), "cuInit")
        st["dev"] = _ck(cu.cuDeviceGet(device), "cuDeviceGet")
        prop = cu.CUmemAllocationProp()
        prop.type = cu.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
        prop.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
        prop.location.id = device
        prop.requestedHandleTypes = H
        mcprop = cu.CUmulticastObjectProp()
        mcprop.numDevices = world
        mcprop.handleTypes = H
        mcprop.flags = 0
        nbytes = numel * _DT_SIZE[dtype]
        mcprop.size = nbytes
        gran_mc = _ck(cu.cuMulticastGetGranularity(mcprop, cu.CUmulticastGranularity_flags.CU_MULTICAST_GRANULARITY_RECOMMENDED),
                      "cuMulticastGetGranularity")
        gran_mem = _ck(cu.cuMemGetAllocationGranularity(prop, cu.CUmemAllocationGranularity_flags.CU_MEM_ALLOC_GRANULARITY_RECOMMENDED),
                       "cuMemGetAllocationGranularity")
        st["gran"] = gran = max(int(gran_mc), int(gran_mem))
        st["size"] = size = (nbytes + gran - 1) // gran * gran
        mcprop.size = size
        st["h_local"] = _ck(cu.cuMemCreate(size, prop, 0), "cuMemCreate")
        st["loose"].append(st["h_local"])
        st["fd_local"] = int(_ck(cu.cuMemExportToShareableHandle(st["h_local"], H, 0), "cuMemExportToShareableHandle"))
        st["fds"].append(st["fd_local"])
        if rank == 0:
            st["mc"] = _ck(cu.cuMulticastCreate(mcprop), "cuMulticastCreate")
            st["loose"].append(st["mc"])
            st["fd_mc"] = int(_ck(cu.cuMemExportToShareableHandle(st["mc"], H, 0), "export multicast handle"))
            st["fds"].append(st["fd_mc"])

    stage(s_create)
    tag = agree(os.getpid() if rank == 0 else None)[0]

    # ---- stage 2: every rank listens on an AF_UNIX socket --------------------------------------------------------
    base = sock_dir or tempfile.gettempdir()
    paths = [os.path.join(base, f"hb-nvls-{tag}-{name.replace('.', '_')}-{r}.sock") for r in range(world)]

    def s_listen() -> None:
        if os.path.exists(paths[rank]):
            os.unlink(paths[rank])
        st["path"] = paths[rank]
        server = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        st["server"] = server
        server.bind(paths[rank])
        server.listen(world)
        server.settimeout(60)                      # a peer that dies mid-way must not hang this rank forever

    stage(s_listen)
    agree()                                        # everyone is listening

    # ---- stage 3: descriptors travel (SCM_RIGHTS), peers' memory is mapped, the device joins the team --------------
    def map_handle(h) -> int:
        size = st["size"]
        va = _ck(cu.cuMemAddressReserve(size, st["gran"], 0, 0), "cuMemAddressReserve")
        try:
            _ck(cu.cuMemMap(va, size, 0, h, 0), "cuMemMap")
        except Exception:
            cu.cuMemAddressFree(va, size)
            raise
        st["mapped"].append((va, size, h))
        if h in st["loose"]:
            st["loose"].remove(h)
        _ck(cu.cuMemSetAccess(va, size, [st["access"]], 1), "cuMemSetAccess")
        return int(va)

    ptrs = [0] * world

    def s_swap() -> None:
        recv_box: dict = {}
        recv_err: list = []

        def receiver() -> None:
            try:
                recv_box.update(_recv_fds(st["server"], world - 1, 2))
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
        st["server"].close()
        st["server"] = None
        os.unlink(paths[rank])
        for fds in recv_box.values():
            st["fds"].extend(fds)
        if send_err or recv_err:
            raise send_err or recv_err[0]

        access = cu.CUmemAccessDesc()
        access.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
        access.location.id = device
        access.flags = cu.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE
        st["access"] = access
        ptrs[rank] = map_handle(st["h_local"])
        for p in range(world):
            if p == rank:
                continue
            hp = _ck(cu.cuMemImportFromShareableHandle(recv_box[p][0], H), "cuMemImportFromShareableHandle")
            st["loose"].append(hp)
            ptrs[p] = map_handle(hp)
        if rank != 0:
            st["mc"] = _ck(cu.cuMemImportFromShareableHandle(recv_box[0][1], H), "import multicast handle")
            st["loose"].append(st["mc"])
        for fd in st["fds"]:
            os.close(fd)
        st["fds"].clear()
        _ck(cu.cuMulticastAddDevice(st["mc"], st["dev"]), "cuMulticastAddDevice")

    stage(s_swap)
    agree()                                        # every device is part of the team before anyone binds

    # ---- stage 4: bind the local memory to the multicast object -------------------------------------------------------
    stage(lambda: _ck(cu.cuMulticastBindMem(st["mc"], 0, st["h_local"], 0, st["size"], 0), "cuMulticastBindMem"))
    agree()

    # ---- stage 5: map the multicast view, zero the local slice ----------------------------------------------------------
    def s_view() -> None:
        st["mc_ptr"] = map_handle(st["mc"])
        local = tensor_from_ptr(ptrs[rank], numel, dtype, device)
        local.zero_()
        torch.cuda.synchronize()
        st["local"] = local

    stage(s_view)
    agree()
    buf = NvlsBuffer(name, numel, dtype, st["local"], ptrs, mc_ptr=st["mc_ptr"], nbytes_mapped=st["size"],
                     _handles=list(st["mapped"]), _h_local=st["h_local"], _gran=st["gran"])
    group.buffers[name] = buf
    return buf
