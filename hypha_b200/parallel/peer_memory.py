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
        self._local_devices = list(local_devices) if local_devices is not None else None
        self.buffers: Dict[str, SymmetricBuffer] = {}
        self.counters = BandwidthCounters()
        self._opened: List[int] = []
        self._owned: List[int] = []

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
