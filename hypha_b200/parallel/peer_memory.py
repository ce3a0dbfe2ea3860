"""Symmetric peer-mapped device buffers for the ranks of one NVSwitch box.

B200-native replacement for the reference's tensor push/pull streams
(crates/network/src/stream_push.rs:16-90, crates/worker/src/connector/mod.rs:318-433):
each rank allocates its buffers once, exports CUDA-IPC handles over the control
plane (``torch.distributed`` object collectives, or any callable transport such
as the loopback bus), maps every peer's buffers, and kernels then read/write
peer HBM directly over NVLink.

Byte counters mirror the reference's ``hypha.bandwidth.{inbound,outbound}.bytes``
instruments (crates/telemetry/src/bandwidth.rs:47-55).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence

import torch

from ..ops.native import load_C
from ..telemetry.bandwidth import BandwidthCounters

_DT_SIZE = {torch.bfloat16: 2, torch.float16: 2, torch.float32: 4, torch.uint8: 1, torch.int32: 4,
            torch.uint32: 4, torch.int64: 8}


class _CudaArray:
    """Minimal ``__cuda_array_interface__`` carrier for a raw device pointer."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {
            "shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3,
            "strides": None,
        }


def tensor_from_ptr(ptr: int, numel: int, dtype: torch.dtype, device: int) -> torch.Tensor:
    nbytes = numel * _DT_SIZE[dtype]
    with torch.cuda.device(device):
        t = torch.as_tensor(_CudaArray(ptr, nbytes), device=torch.device("cuda", device))
    return t.view(dtype)


@dataclass
class SymmetricBuffer:
    """A buffer that exists on every rank; ``peer_ptrs[r]`` is rank r's copy as mapped here."""

    name: str
    numel: int
    dtype: torch.dtype
    local: torch.Tensor
    peer_ptrs: List[int] = field(default_factory=list)

    @property
    def nbytes(self) -> int:
        return self.numel * _DT_SIZE[self.dtype]


def store_exchange(store, rank: int, world: int, prefix: str = "hb-xchg", timeout_s: float = 120.0) -> Callable[[object], List[object]]:
    """Control-plane all-gather over a key-value store (``torch.distributed.TCPStore`` or anything with
    ``set`` / ``get`` / ``wait``): needs no process group, so a membership generation that contains a
    replacement PROCESS can use it (a torch process group cannot admit one).  Every call of the returned
    function is one round; all ``world`` ranks must make the same sequence of calls."""
    import base64
    import json
    from datetime import timedelta

    seq = [0]

    def enc(o):
        # JSON with two tagged extensions (bytes for IPC handles, int-keyed dicts for owner maps): values read
        # from the store are never unpickled (the TCPStore is unauthenticated)
        if isinstance(o, (bytes, bytearray)):
            return {"__b64__": base64.b64encode(bytes(o)).decode()}
        if isinstance(o, dict):
            if all(isinstance(k, str) for k in o):
                return {k: enc(v) for k, v in o.items()}
            return {"__items__": [[enc(k), enc(v)] for k, v in o.items()]}
        if isinstance(o, (list, tuple)):
            return [enc(v) for v in o]
        return o

    def dec(o):
        if isinstance(o, dict):
            if "__b64__" in o and len(o) == 1:
                return base64.b64decode(o["__b64__"])
            if "__items__" in o and len(o) == 1:
                return {(tuple(k) if isinstance(k, list) else k): v for k, v in ((dec(a), dec(b)) for a, b in o["__items__"])}
            return {k: dec(v) for k, v in o.items()}
        if isinstance(o, list):
            return [dec(v) for v in o]
        return o

    def exchange(obj):
        n = seq[0]
        seq[0] += 1
        store.set(f"{prefix}/{n}/{rank}", json.dumps(enc(obj)))
        keys = [f"{prefix}/{n}/{r}" for r in range(world)]
        store.wait(keys, timedelta(seconds=timeout_s))
        return [dec(json.loads(store.get(k))) for k in keys]

    return exchange


class RebindError(RuntimeError):
    pass


class PeerGroup:
    """Ranks of one box that share symmetric buffers.

    ``exchange(obj) -> list[obj]`` is the control-plane all-gather used to swap
    IPC handles.  ``PeerGroup.from_torch_distributed()`` uses the default process
    group; ``PeerGroup.single_process(devices)`` drives several GPUs from one
    process through direct peer access (tests / notebooks)."""

    def __init__(self, rank: int, world: int, device: int,
                 exchange: Optional[Callable[[object], List[object]]] = None,
                 local_devices: Optional[Sequence[int]] = None):
        self.rank, self.world, self.device = rank, world, device
        self._exchange = exchange
        self.deferred = False        # a joining process allocates locally first and maps peers in rebind()
        self._local_devices = list(local_devices) if local_devices is not None else None
        self.buffers: Dict[str, SymmetricBuffer] = {}
        self.counters = BandwidthCounters()
        self._opened: List[int] = []
        self._owned: List[int] = []
        self._vmm_peer_maps: List[tuple] = []     # VMM mappings of survivors' NVLS buffers (replacement process)
        self.multicast_broken = False             # a member of the multicast team was replaced: P2P only

    # -- constructors ---------------------------------------------------------------------
    @classmethod
    def from_torch_distributed(cls, device: Optional[int] = None) -> "PeerGroup":
        import torch.distributed as dist

        rank, world = dist.get_rank(), dist.get_world_size()
        device = torch.cuda.current_device() if device is None else device

        def exchange(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out

        return cls(rank, world, device, exchange)

    @classmethod
    def solo(cls, device: Optional[int] = None) -> "PeerGroup":
        device = torch.cuda.current_device() if device is None else device
        return cls(0, 1, device, lambda obj: [obj])

    # -- allocation -----------------------------------------------------------------------
    def alloc(self, name: str, numel: int, dtype: torch.dtype) -> SymmetricBuffer:
        C = load_C()
        nbytes = numel * _DT_SIZE[dtype]
        with torch.cuda.device(self.device):
            ptr = C.sym_alloc(nbytes, True)
        self._owned.append(ptr)
        local = tensor_from_ptr(ptr, numel, dtype, self.device)
        buf = SymmetricBuffer(name, numel, dtype, local)
        if self.world == 1:
            buf.peer_ptrs = [ptr]
        elif self.deferred:
            buf.peer_ptrs = [ptr if r == self.rank else 0 for r in range(self.world)]
        else:
            handles = self._exchange((self.rank, C.ipc_export(ptr)))
            ptrs = [0] * self.world
            for r, h in handles:
                if r == self.rank:
                    ptrs[r] = ptr
                else:
                    with torch.cuda.device(self.device):
                        p = C.ipc_open(h)
                    self._opened.append(p)
                    ptrs[r] = p
            buf.peer_ptrs = ptrs
        self.buffers[name] = buf
        return buf

    @classmethod
    def joining(cls, rank: int, world: int, device: int) -> "PeerGroup":
        """The group handle of a replacement process: ``alloc`` creates the local buffers only (the survivors
        are not allocating anything, so there is nobody to exchange with); ``rebind`` then maps the peers."""
        g = cls(rank, world, device, None)
        g.deferred = True
        return g

    def rebind(self, exchange: Callable[[object], List[object]], replaced: Sequence[int],
               members: Optional[Sequence[int]] = None) -> None:
        """Collective over the processes of a NEW membership generation (same rank numbering, ``replaced`` =
        ranks whose process is new): every rank publishes IPC handles of all its buffers; survivors drop
        their mappings of the dead processes' buffers and open the newcomers', newcomers open everybody's.
        Mappings between two survivors are left alone.  Kernels pick the new addresses up at their next
        launch (pointer tables are read from ``SymmetricBuffer.peer_ptrs``).  ``members`` = the ranks that
        take part (default: all ``world``; ranks whose slot is vacant are skipped).  Survivors' NVLS buffers (VMM
        allocations: no CUDA-IPC handle exists for them) reach a newcomer as POSIX file descriptors over AF_UNIX
        and are mapped there for plain P2P access; the multicast team itself is fixed at creation, so the group
        runs the P2P variant of the outer step from then on (``multicast_broken``)."""
        C = load_C()
        replaced = sorted(set(int(r) for r in replaced))
        self._exchange = exchange
        names = sorted(self.buffers)
        i_am_new = self.rank in replaced

        def describe(n):
            buf = self.buffers[n]
            if getattr(buf, "mc_ptr", 0) or getattr(buf, "_h_local", None) is not None:
                # VMM allocation (NVLS buffer): no CUDA-IPC handle exists for it; newcomers get a file descriptor
                return (buf.numel, str(buf.dtype), "vmm", int(getattr(buf, "nbytes_mapped", 0)))
            return (buf.numel, str(buf.dtype), C.ipc_export(buf.local.data_ptr()), 0)

        mine = {n: describe(n) for n in names}
        table = {int(r): m for r, m in exchange((self.rank, mine))}
        expect = sorted(int(m) for m in members) if members is not None else list(range(self.world))
        if sorted(table) != expect:
            raise RebindError(f"membership exchange returned ranks {sorted(table)}, expected {expect}")
        problems = []
        for r, bufs in table.items():
            if sorted(bufs) != names:
                problems.append(f"rank {r} holds buffers {sorted(bufs)}, this rank {names}")
                continue
            for n in names:
                if tuple(bufs[n][:2]) != (self.buffers[n].numel, str(self.buffers[n].dtype)):
                    problems.append(f"rank {r} buffer {n!r} is {bufs[n][:2]}, here {(self.buffers[n].numel, str(self.buffers[n].dtype))}")
        # agree on the verdict before anybody changes a mapping
        verdicts = exchange(problems[:3])
        bad = [p for v in verdicts for p in v]
        if bad:
            raise RebindError("; ".join(bad[:3]))
        # ---- VMM (NVLS) buffers of survivors reach the newcomers as file descriptors over AF_UNIX -----------------
        survivors = [r for r in expect if r not in replaced]
        vmm_names = [n for n in names if any(table[r][n][2] == "vmm" for r in survivors)]
        fd_maps: Dict[int, Dict[str, int]] = {}          # on a newcomer: survivor rank -> {buffer name: fd}
        if vmm_names and replaced:
            import os
            import socket
            import tempfile
            import threading

            from . import nvls as _nvls

            tag = exchange(os.getpid() if self.rank == expect[0] else None)[0]
            paths = {r: os.path.join(tempfile.gettempdir(), f"hb-rebind-{tag}-{r}.sock") for r in replaced if r in expect}
            server = None
            if i_am_new:
                if os.path.exists(paths[self.rank]):
                    os.unlink(paths[self.rank])
                server = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
                server.bind(paths[self.rank])
                server.listen(len(expect))
                server.settimeout(60)
            exchange("listening")
            err = None
            try:
                if i_am_new:
                    got = _nvls._recv_fds(server, len(survivors), len(vmm_names))
                    for r, fds in got.items():
                        fd_maps[int(r)] = dict(zip(vmm_names, fds))
                else:
                    fds = [self.buffers[n].export_fd() for n in vmm_names]
                    try:
                        for r in paths:
                            _nvls._send_fds(paths[r], self.rank, fds)
                    finally:
                        for fd in fds:
                            os.close(fd)
            except Exception as exc:                      # noqa: BLE001 - agreed on below
                err = f"rank {self.rank}: {type(exc).__name__}: {exc}"
            finally:
                if server is not None:
                    server.close()
                    if os.path.exists(paths[self.rank]):
                        os.unlink(paths[self.rank])
            errs = [e for e in exchange(err) if e]
            if errs:
                raise RebindError("; ".join(errs[:3]))
        for r in expect:
            if r == self.rank or not (i_am_new or r in replaced):
                continue
            for n in names:
                buf = self.buffers[n]
                old = buf.peer_ptrs[r] if r < len(buf.peer_ptrs) else 0
                if old and old in self._opened:
                    try:
                        C.ipc_close(old)            # the exporting process is gone; the mapping is stale
                    except Exception:
                        pass
                    self._opened.remove(old)
                kind = table[r][n][2]
                if kind == "vmm":
                    from . import nvls as _nvls

                    with torch.cuda.device(self.device):
                        p, mapping = _nvls.map_peer_fd(fd_maps[r][n], int(table[r][n][3]), self.device)
                    import os

                    os.close(fd_maps[r][n])
                    self._vmm_peer_maps.append(mapping)
                else:
                    with torch.cuda.device(self.device):
                        p = C.ipc_open(kind)
                    self._opened.append(p)
                buf.peer_ptrs[r] = p
        if vmm_names and replaced:
            # the NVSwitch multicast team is fixed at creation: with a replaced member the multimem path is off
            # for good (plain P2P loads/stores go on working); DiLoCoOuter checks this flag
            self.multicast_broken = True
        self.deferred = False
        exchange(self.rank)                           # nobody launches before every mapping exists

    def pull(self, src_rank: int, name: str, num_ctas: int = 0) -> None:
        """Copy rank ``src_rank``'s copy of ``name`` into the local copy over NVLink
        (parameter broadcast for a (re)joining worker)."""
        buf = self.buffers[name]
        from ..ops.native import cuda_stream

        load_C().peer_pull_copy(buf.peer_ptrs[src_rank], buf.local.data_ptr(), buf.nbytes, num_ctas,
                                cuda_stream())
        self.counters.add("/hypha-peer-memory/pull", inbound=buf.nbytes)

    def close(self) -> None:
        C = load_C()
        for p in self._opened:
            try:
                C.ipc_close(p)
            except Exception:
                pass
        self._opened.clear()
        if self._vmm_peer_maps:
            from . import nvls as _nvls

            for mp in self._vmm_peer_maps:
                _nvls.unmap_peer(mp)
            self._vmm_peer_maps.clear()
        for b in self.buffers.values():
            if hasattr(b, "close"):          # NVLS (VMM + multicast) buffers release their own mappings
                b.close()
        self.buffers.clear()
        for p in self._owned:
            try:
                C.sym_free(p)
            except Exception:
                pass
        self._owned.clear()


def single_process_groups(devices: Sequence[int]) -> List[PeerGroup]:
    """One ``PeerGroup`` per device, all in this process, wired through peer access.

    ``alloc`` must then be called via ``alloc_all(groups, ...)`` so every group sees
    every pointer."""
    C = load_C()
    for d in devices:
        for e in devices:
            if d != e and not C.enable_peer_access(d, e):
                raise RuntimeError(f"no peer access {d}->{e}")
    return [PeerGroup(r, len(devices), d, None, devices) for r, d in enumerate(devices)]


def alloc_all(groups: Sequence[PeerGroup], name: str, numel: int, dtype: torch.dtype) -> List[SymmetricBuffer]:
    C = load_C()
    ptrs = []
    for g in groups:
        with torch.cuda.device(g.device):
            p = C.sym_alloc(numel * _DT_SIZE[dtype], True)
        g._owned.append(p)
        ptrs.append(p)
    out = []
    for g, p in zip(groups, ptrs):
        buf = SymmetricBuffer(name, numel, dtype, tensor_from_ptr(p, numel, dtype, g.device), list(ptrs))
        g.buffers[name] = buf
        out.append(buf)
    return out
