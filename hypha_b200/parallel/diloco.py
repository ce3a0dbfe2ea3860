"""DiLoCo on B200: fused flat AdamW inner optimizer + fused NVLink outer step.

Algorithm (reference: docs/training.md:18-25,268-320; executor loop
executors/accelerate/.../training.py:105-151; parameter server
crates/worker/src/executor/parameter_server.rs:177-283,386-446):

    repeat rounds:
        H inner AdamW steps on every worker            (no communication)
        delta_r = theta_r - theta_outer                 (pseudo-gradient, sign as utils.py:118-123)
        g       = mean_r delta_r                        (true 1/N mean; ref quirk Q1 fixed)
        m       = g            if first round else mu*m + g
        theta_outer += lr * (mu*m + g)                  (Nesterov)
        every worker restarts from theta_outer

Three interchangeable outer back-ends:
    "fused"  one sm_100a kernel per rank over peer memory (product path, no NCCL)
    "nccl"   dist.all_reduce(delta) + unfused torch ops (the baseline to beat)
    "local"  single worker (N=1), exact fp32
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional

import torch

from ..ops.native import cuda_stream, load_C
from .peer_memory import PeerGroup, SymmetricBuffer


# -----------------------------------------------------------------------------------
# configs
# -----------------------------------------------------------------------------------
@dataclass
class AdamWConfig:
    lr: float = 1e-4
    betas: tuple = (0.9, 0.999)
    eps: float = 1e-8
    weight_decay: float = 0.01          # torch.optim.AdamW default, as in the reference
    state_dtype: str = "bf16"           # "bf16" (stochastic rounding) or "fp32"


@dataclass
class DiLoCoConfig:
    inner_steps: int = 100              # H
    outer_lr: float = 0.7               # scheduler_config.rs:89-92 defaults
    outer_momentum: float = 0.9
    backend: str = "auto"               # auto | fused | sharded | nccl | local
    num_ctas: int = 0                   # 0 = one CTA per SM
    pipeline_chunks: int = 0            # fused backend: 0 = auto (by size), 1 = plain A|B|C, K = K-deep pipeline
    link_ctas: int = 0                  # CTAs driving NVLink while the others stream HBM (0 = auto)
    timeout_ms: int = 30000


# -----------------------------------------------------------------------------------
# LR schedules (reference: utils.py:90-102 -> transformers.optimization)
# -----------------------------------------------------------------------------------
def make_lr_lambda(cfg: Optional[dict]) -> Callable[[int], float]:
    """Multiplier(step) for constant / cosine-with-warmup / linear-with-warmup / wsd.

    Field names follow the wire format (crates/messages/src/lib.rs:671-687); the
    reference's ``decay_step`` typo (quirk Q4) is accepted as an alias."""
    if not cfg or not cfg.get("type"):
        return lambda step: 1.0
    kind = cfg["type"]
    warm = int(cfg.get("warmup_steps", 0))
    if kind == "cosine-with-warmup":
        total = int(cfg["training_steps"])

        def f(step):
            if step < warm:
                return step / max(1, warm)
            prog = (step - warm) / max(1, total - warm)
            return max(0.0, 0.5 * (1.0 + math.cos(math.pi * min(1.0, prog))))
        return f
    if kind == "linear-with-warmup":
        total = int(cfg["training_steps"])

        def f(step):
            if step < warm:
                return step / max(1, warm)
            return max(0.0, (total - step) / max(1, total - warm))
        return f
    if kind == "wsd":
        # warmup -> stable -> cosine decay (transformers.get_wsd_schedule semantics).  The stable
        # length comes from `stable_steps`, or `training_steps - warmup - decay`, else 0.
        decay = int(cfg.get("decay_steps", cfg.get("decay_step", 0)))
        if "stable_steps" in cfg:
            stable = int(cfg["stable_steps"])
        elif "training_steps" in cfg:
            stable = max(0, int(cfg["training_steps"]) - warm - decay)
        else:
            stable = 0

        def f(step):
            if step < warm:
                return step / max(1, warm)
            if step < warm + stable:
                return 1.0
            if decay <= 0:
                return 1.0
            prog = min(1.0, (step - warm - stable) / decay)
            return max(0.0, 0.5 * (1.0 + math.cos(math.pi * prog)))
        return f
    raise ValueError(f"unknown LR scheduler type {kind!r}")


# -----------------------------------------------------------------------------------
# inner optimizer
# -----------------------------------------------------------------------------------
class FlatAdamW:
    """AdamW over a flat bf16 parameter/gradient buffer with fp32 master weights.

    One kernel launch per step (``adamw_flat``); semantics = ``torch.optim.AdamW``
    (reference: utils.py:56-64)."""

    def __init__(self, params: torch.Tensor, grads: torch.Tensor, cfg: AdamWConfig,
                 lr_lambda: Optional[Callable[[int], float]] = None):
        assert params.dtype == torch.bfloat16 and params.numel() % 8 == 0
        self.cfg = cfg
        self.params, self.grads = params, grads
        self.master = params.float()
        sdt = torch.bfloat16 if cfg.state_dtype == "bf16" else torch.float32
        self.m = torch.zeros(params.numel(), dtype=sdt, device=params.device)
        self.v = torch.zeros(params.numel(), dtype=sdt, device=params.device)
        self.step_count = 0
        self.lr_lambda = lr_lambda or (lambda s: 1.0)

    @property
    def lr(self) -> float:
        return self.cfg.lr * self.lr_lambda(self.step_count)

    def begin_step(self) -> None:
        """Advance the step counter and fix this step's scalars (lr, bias corrections, SR seed)."""
        c = self.cfg
        self._lr = self.lr
        self.step_count += 1
        t = self.step_count
        self._bc1 = 1.0 - c.betas[0] ** t
        self._bc2 = 1.0 - c.betas[1] ** t
        self._seed = (t * 2654435761) & 0xFFFFFFFF

    def step_range(self, begin: int, end: int, grad_scale: float = 1.0) -> None:
        """AdamW on the flat slice ``[begin, end)`` (multiples of 8) on the CURRENT stream.  Lets the
        trainer update a layer as soon as its gradients exist, overlapped with the rest of backward."""
        c = self.cfg
        n = end - begin
        if n <= 0:
            return
        assert begin % 8 == 0 and n % 8 == 0
        es = 2 if c.state_dtype == "bf16" else 4
        load_C().adamw_flat(self.params.data_ptr() + 2 * begin, self.grads.data_ptr() + 2 * begin,
                            self.master.data_ptr() + 4 * begin, self.m.data_ptr() + es * begin,
                            self.v.data_ptr() + es * begin, int(c.state_dtype == "bf16"), n, self._lr,
                            c.betas[0], c.betas[1], c.eps, c.weight_decay, self._bc1, self._bc2, grad_scale,
                            (self._seed + begin) & 0xFFFFFFFF, cuda_stream())

    def finish_step(self) -> None:
        from ..ops.functional import bump_fp8_weight_version

        bump_fp8_weight_version()           # cached fp8 weight copies are stale now

    def step(self, grad_scale: float = 1.0) -> None:
        self.begin_step()
        self.step_range(0, self.params.numel(), grad_scale)
        self.finish_step()

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return {"master": self.master, "m": self.m, "v": self.v,
                "step": torch.tensor(self.step_count)}

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        self.master.copy_(sd["master"])
        self.m.copy_(sd["m"].to(self.m.dtype))
        self.v.copy_(sd["v"].to(self.v.dtype))
        self.step_count = int(sd["step"])


# -----------------------------------------------------------------------------------
# outer optimizer
# -----------------------------------------------------------------------------------
def shard_bounds(n: int, world: int, rank: int, align: int = 1024):
    per = (n // world + align - 1) // align * align
    b = min(n, rank * per)
    e = min(n, b + per) if rank < world - 1 else n
    return b, max(0, e - b)


class _PyEpochTable:
    """Plain-Python stand-in for the native epoch table when the group spans more than one box (world > 8:
    NCCL backend, no fused kernel launches to keep consistent)."""

    def __init__(self, world: int):
        self._s = {"epoch": 0, "flag_epoch": 0, "mask": (1 << world) - 1, "world": world, "owner": list(range(world))}

    def snapshot(self) -> dict:
        return dict(self._s, owner=list(self._s["owner"]))

    def _bump(self) -> int:
        self._s["epoch"] += 1
        return self._s["epoch"]

    def set_mask(self, mask: int) -> int:
        self._s["mask"] = int(mask)
        return self._bump()

    def set_owner(self, shard: int, rank: int) -> int:
        self._s["owner"][shard] = int(rank)
        return self._bump()

    def set_flag_epoch(self, v: int) -> int:
        self._s["flag_epoch"] = int(v)
        return self._bump()


class _OwnerMap:
    """``{shard: rank}`` view of the epoch table's ownership column with the dict operations the membership code
    uses (every rank applies the same events in the same order, so the tables agree without communication)."""

    def __init__(self, table, world: int):
        self._t, self._world = table, world

    def _col(self) -> List[int]:
        return [int(x) for x in self._t.snapshot()["owner"]]

    def __getitem__(self, k: int) -> int:
        return self._col()[int(k)]

    def __setitem__(self, k: int, v: int) -> None:
        self._t.set_owner(int(k), int(v))

    def get(self, k, default=None):
        k = int(k)
        return self._col()[k] if 0 <= k < self._world else default

    def keys(self):
        return range(self._world)

    def values(self):
        return self._col()

    def items(self):
        return list(enumerate(self._col()))

    def __iter__(self):
        return iter(range(self._world))

    def __len__(self) -> int:
        return self._world

    def __eq__(self, other) -> bool:
        return dict(self.items()) == dict(other.items()) if hasattr(other, "items") else NotImplemented

    def __repr__(self) -> str:
        return repr(dict(self.items()))


class DiLoCoOuter:
    """Outer Nesterov optimizer, ZeRO-1 sharded across the participating ranks."""

    def __init__(self, opt: FlatAdamW, cfg: DiLoCoConfig, group: Optional[PeerGroup] = None,
                 param_buf: Optional[SymmetricBuffer] = None,
                 delta_buf: Optional[SymmetricBuffer] = None, process_group=None):
        self.opt, self.cfg = opt, cfg
        self.group = group
        self.process_group = process_group
        self.n = opt.params.numel()
        dev = opt.params.device
        world = group.world if group is not None else 1
        backend = cfg.backend
        if world == 1 and process_group is not None and backend in ("auto", "sharded"):
            import torch.distributed as dist

            world = dist.get_world_size(process_group)        # a team without peer memory: several boxes
        if backend == "auto":
            if world == 1:
                backend = "local"
            elif param_buf is not None and world <= 8:
                backend = "fused"
            else:
                # across boxes there is no peer memory: stock collectives, outer state sharded like the fused path
                import torch.distributed as dist

                backend = "sharded" if (world > 8 and dist.is_available() and dist.is_initialized()) else "nccl"
        self.backend = backend
        self.world, self.rank = world, (group.rank if group is not None else 0)
        if backend == "sharded":
            import torch.distributed as dist

            self.world, self.rank = dist.get_world_size(process_group), dist.get_rank(process_group)
            world = self.world
        self.round = 0
        # membership, shard ownership and the flag-token epoch live in the native epoch table (`csrc/comm/vmm.cpp`:
        # writers publish under a sequence lock, a launch reads a consistent snapshot); one box = 8 ranks at most
        self._table.snapshot()
        self.last_status = (0, 0)
        if backend == "local":
            self.shard_begin, self.shard_len = 0, self.n
            self.theta_outer = opt.master.clone()
            self.momentum = torch.zeros(self.n, dtype=torch.float32, device=dev)
        elif backend == "fused":
            if world > 8:
                raise ValueError(f"the fused outer step covers one NVSwitch box (<= 8 ranks); got world = {world}: use "
                                 f"backend='sharded' (or 'auto') across boxes")
            assert group is not None and param_buf is not None and delta_buf is not None
            assert param_buf.local.data_ptr() == opt.params.data_ptr(), \
                "fused outer step needs the model parameters in the symmetric buffer"
            self.param_buf, self.delta_buf = param_buf, delta_buf
            self.anchor_buf = group.alloc("diloco.anchor", self.n, torch.bfloat16)
            self.flag_buf = group.alloc("diloco.flags", 64, torch.int32)
            self.anchor_buf.local.copy_(opt.params)
            self.shard_begin, self.shard_len = shard_bounds(self.n, world, self.rank)
            sl = slice(self.shard_begin, self.shard_begin + self.shard_len)
            self.theta_outer = opt.master[sl].clone()
            self.momentum = torch.zeros(self.shard_len, dtype=torch.float32, device=dev)
            # segments owned by this rank: its shard + shards adopted from dropped ranks
            self.segments = {self.rank: (self.shard_begin, self.shard_len, self.theta_outer, self.momentum)}
            self.layout = [shard_bounds(self.n, world, r) for r in range(world)]
            # multimem path when the buffers carry a multicast address; measured slower than P2P at
            # N=2 (nothing to save: one peer) and 12-19 % faster at N=8
            self.use_nvls = world >= 4
            load_C().set_outer_timeout_ms(int(cfg.timeout_ms))
        elif backend == "sharded":
            # more than one NVSwitch box (or no peer memory): reduce-scatter / all-gather of the process group, with
            # the fp32 outer state sharded over ALL ranks as on the fused path (8 bytes x n / world per rank instead
            # of the 8 bytes x n of the replicated "nccl" baseline) and the same numerics: the round starts from the
            # bf16 weights every rank holds (the anchor), theta_outer / momentum stay fp32 on the shard's owner
            self._per = (self.n + world - 1) // world
            self._per = (self._per + 1023) // 1024 * 1024           # equal shards; the tail past n is padding
            self.shard_begin = min(self.n, self.rank * self._per)
            self.shard_len = max(0, min(self.n, self.shard_begin + self._per) - self.shard_begin)
            sl = slice(self.shard_begin, self.shard_begin + self.shard_len)
            self.theta_outer = opt.master[sl].clone()
            self.momentum = torch.zeros(self.shard_len, dtype=torch.float32, device=dev)
            self.anchor = opt.params.clone()
        elif backend == "nccl":
            # baseline: replicated outer state, stock collectives + unfused torch ops
            self.shard_begin, self.shard_len = 0, self.n
            self.theta_outer = opt.master.clone()
            self.momentum = torch.zeros(self.n, dtype=torch.float32, device=dev)
        else:
            raise ValueError(backend)

    # -- membership (fault tolerance: SURVEY 5.3, BASELINE config 4) ------------------------
    @property
    def _table(self):
        t = self.__dict__.get("_table_obj")
        if t is None:
            world = self.world
            t = self.__dict__["_table_obj"] = load_C().EpochTable(world) if world <= 8 else _PyEpochTable(world)
            self.__dict__["_owner_map"] = _OwnerMap(t, world)
        return t

    @property
    def member_mask(self) -> int:
        return int(self._table.snapshot()["mask"])

    @member_mask.setter
    def member_mask(self, mask: int) -> None:
        self._table.set_mask(int(mask))

    @property
    def attempt(self) -> int:
        """Flag-token epoch of the device-side barriers (bumped by every launch)."""
        return int(self._table.snapshot()["flag_epoch"])

    @attempt.setter
    def attempt(self, v: int) -> None:
        self._table.set_flag_epoch(int(v))

    @property
    def _owner(self) -> "_OwnerMap":
        self._table
        return self.__dict__["_owner_map"]

    @_owner.setter
    def _owner(self, value) -> None:
        om = self._owner                            # a plain {shard: rank} map agreed over the control plane
        for k, v in value.items():
            om[int(k)] = int(v)

    def set_members(self, ranks: List[int]) -> None:
        mask = 0
        for r in ranks:
            mask |= 1 << r
        self.member_mask = mask

    def step(self) -> None:
        """Run one outer step; all participating ranks must call it."""
        opt, cfg = self.opt, self.cfg
        first = int(self.round == 0)
        if self.backend in ("local", "fused"):         # the collective backends are stream-ordered torch ops
            C = load_C()
            s = cuda_stream()
        if self.backend == "local":
            C.diloco_outer_local(opt.master.data_ptr(), 0, 0, self.theta_outer.data_ptr(),
                                 self.momentum.data_ptr(), opt.master.data_ptr(),
                                 opt.params.data_ptr(), self.n, cfg.outer_lr, cfg.outer_momentum,
                                 first, 1.0, s)
        elif self.backend == "fused":
            self.attempt += 1                      # flag tokens must be unique per launch
            snap = self._table.snapshot()          # ONE consistent (members, flag epoch) pair for this launch
            segs = [self.segments[k] for k in sorted(self.segments)]
            C.diloco_outer_step_fused(self.delta_buf.peer_ptrs, self.param_buf.peer_ptrs,
                                      self.anchor_buf.peer_ptrs, self.flag_buf.peer_ptrs,
                                      opt.master.data_ptr(), [b for b, _, _, _ in segs],
                                      [l for _, l, _, _ in segs], [t.data_ptr() for _, _, t, _ in segs],
                                      [m.data_ptr() for _, _, _, m in segs], self.n, self.rank,
                                      int(snap["mask"]), int(snap["flag_epoch"]), cfg.outer_lr, cfg.outer_momentum,
                                      first, cfg.num_ctas, *self._multicast_ptrs(),
                                      [b for b, _ in self.layout], [l for _, l in self.layout],
                                      *self._pipeline_shape(), s)
            nmem = bin(int(snap["mask"])).count("1")
            wire = 2 * sum(l for _, l, _, _ in segs) * (nmem - 1)
            self.group.counters.add("/hypha-peer-memory/outer-step", inbound=wire, outbound=wire)
        elif self.backend == "sharded":
            self._step_sharded(bool(first))
        else:
            import torch.distributed as dist

            CH = 1 << 28                       # chunked so the unfused baseline fits next to an 8B model
            delta = torch.empty(self.n, dtype=torch.bfloat16, device=opt.master.device)
            for a in range(0, self.n, CH):
                b = min(self.n, a + CH)
                delta[a:b] = (opt.master[a:b] - self.theta_outer[a:b]).to(torch.bfloat16)
            dist.all_reduce(delta, group=self.process_group)
            for a in range(0, self.n, CH):
                b = min(self.n, a + CH)
                g = delta[a:b].float().div_(self.world)
                if first:
                    self.momentum[a:b].copy_(g)
                else:
                    self.momentum[a:b].mul_(cfg.outer_momentum).add_(g)
                self.theta_outer[a:b].add_((self.momentum[a:b] * cfg.outer_momentum).add_(g).mul_(cfg.outer_lr))
                opt.master[a:b].copy_(self.theta_outer[a:b])
                opt.params[a:b].copy_(self.theta_outer[a:b])
            del delta
        self.round += 1
        from ..ops.functional import bump_fp8_weight_version

        bump_fp8_weight_version()

    def _step_sharded(self, first: bool) -> None:
        """The outer step over stock collectives with ZeRO-1 sharded outer state (the multi-box path).

        Per piece of every shard: delta = master - anchor (bf16 on the wire) -> ``reduce_scatter`` -> the owner's
        Nesterov update in fp32 -> new bf16 weights -> ``all_gather`` -> parameters, master weights and the anchor
        of every rank.  Staging is bounded (two buffers of <= 2^28 elements), so it fits next to an 8B model."""
        import torch.distributed as dist

        opt, cfg, W, pg, n, per = self.opt, self.cfg, self.world, self.process_group, self.n, self._per
        dev = opt.master.device
        # gloo (the CPU plumbing configuration) reduces fp32; NCCL carries the pseudo-gradient in bf16 like the kernel
        wire = torch.bfloat16 if dev.type == "cuda" else torch.float32
        piece = min(per, max(1024, ((1 << 28) // W) // 1024 * 1024))
        send = torch.empty(W * piece, dtype=wire, device=dev)
        mine = torch.empty(piece, dtype=wire, device=dev)
        new = torch.empty(piece, dtype=torch.bfloat16, device=dev)
        gathered = torch.empty(W * piece, dtype=torch.bfloat16, device=dev)
        for off in range(0, per, piece):
            ln = min(piece, per - off)
            spans = []
            for r in range(W):
                a = min(n, r * per + off)
                b = min(n, a + ln, (r + 1) * per)
                spans.append((a, b))
                view = send[r * ln:(r + 1) * ln]
                if b > a:
                    view[:b - a] = (opt.master[a:b] - self.anchor[a:b].float()).to(wire)
                view[max(0, b - a):].zero_()
            try:
                dist.reduce_scatter_tensor(mine[:ln], send[:W * ln], group=pg)
            except (RuntimeError, NotImplementedError):        # gloo has no reduce-scatter: all-reduce, keep a slice
                dist.all_reduce(send[:W * ln], group=pg)
                mine[:ln].copy_(send[self.rank * ln:(self.rank + 1) * ln])
            a, b = spans[self.rank]
            k = b - a
            if k > 0:
                lo = a - self.shard_begin
                g = mine[:k].float().div_(W)
                m = self.momentum[lo:lo + k]
                if first:
                    m.copy_(g)
                else:
                    m.mul_(cfg.outer_momentum).add_(g)
                self.theta_outer[lo:lo + k].add_((m * cfg.outer_momentum).add_(g).mul_(cfg.outer_lr))
                new[:k].copy_(self.theta_outer[lo:lo + k])
            new[max(0, k):ln].zero_()
            dist.all_gather_into_tensor(gathered[:W * ln], new[:ln], group=pg)
            for r, (a, b) in enumerate(spans):
                if b > a:
                    fresh = gathered[r * ln:r * ln + (b - a)]
                    opt.params[a:b].copy_(fresh)
                    self.anchor[a:b].copy_(fresh)
                    opt.master[a:b].copy_(fresh)               # every replica starts the round from the same bf16 weights

    # -- fault tolerance ------------------------------------------------------------------------
    def members(self) -> List[int]:
        return [r for r in range(self.world) if (self.member_mask >> r) & 1]

    def adopter_of(self, dead: int, members: List[int]) -> int:
        """Deterministic adopter of an orphaned shard: the next live rank in cyclic order."""
        for d in range(1, self.world + 1):
            c = (dead + d) % self.world
            if c in members:
                return c
        raise RuntimeError("no live rank left")

    def owners(self) -> Dict[int, int]:
        """shard -> rank that currently runs its outer update.  Replicated bookkeeping: every rank applies the
        same membership events in the same order, so the maps agree without communication."""
        return self._owner

    def drop_ranks(self, dead: List[int]) -> None:
        """Shrink the membership; orphaned shards are adopted with theta_outer re-derived from the
        common bf16 anchor (every rank holds it) and zero momentum.  "Orphaned" includes shards a now-dead
        rank had adopted earlier (cascaded failures): every shard whose owner is not live gets a new one."""
        live = [r for r in self.members() if r not in dead]
        self.set_members(live)
        own = self.owners()
        for k in range(self.world):
            if own[k] in live:
                continue
            own[k] = self.adopter_of(k, live)
            if own[k] == self.rank and k not in self.segments:
                b, l = shard_bounds(self.n, self.world, k)
                theta = self.anchor_buf.local[b:b + l].float()
                self.segments[k] = (b, l, theta, torch.zeros_like(theta))

    def handback(self, exchange, ranks: List[int]) -> None:
        """Owners take their outer state back (collective over every rank of the new membership, returning /
        joining ranks included; ``exchange`` = a control-plane all-gather that reaches all of them).  For each
        rank in ``ranks`` whose shard somebody else has been running, the holder stages its fp32 theta_outer
        segment, then its momentum segment, in its copy of the symmetric delta buffer (free at a round boundary:
        it aliases the gradient buffer) and the owner pulls them over NVLink with the same peer-memory copy
        kernel that re-seeds the parameters.  Nothing is re-derived from bf16 and no momentum is reset: the
        shard continues exactly where the holder left it."""
        own = self.owners()
        as_f32 = self.delta_buf.local.view(torch.float32)
        for r in sorted(int(x) for x in ranks):
            holder = own.get(r, r)
            if holder == r:
                continue
            b, l = shard_bounds(self.n, self.world, r)
            assert l <= as_f32.numel(), "shard does not fit the staging buffer"
            for which in (2, 3):                              # segment tuple index: theta_outer, momentum
                if self.rank == holder:
                    as_f32[:l].copy_(self.segments[r][which])
                    torch.cuda.synchronize()
                exchange(("staged", r, which))
                if self.rank == r:
                    dst = self.theta_outer if which == 2 else self.momentum
                    load_C().peer_pull_copy(self.delta_buf.peer_ptrs[holder], dst.data_ptr(), 4 * l, 0, cuda_stream())
                    torch.cuda.synchronize()
                    self.group.counters.add("/hypha-peer-memory/handback", inbound=4 * l)
                exchange(("pulled", r, which))

    def readmit(self, rank: int, handback: bool = True) -> None:
        """A (re)joined rank takes its shard back at a round boundary (collective: every rank of the new
        membership calls it).  With ``handback`` the holder's fp32 theta_outer / momentum of that shard travel
        back to the owner first (`handback`); the segment is dropped by the holder afterwards.  The returning rank
        missed the membership events of its absence, so the owner map is re-agreed first."""
        if handback and self.backend == "fused" and self.group is not None and self.group._exchange is not None:
            returning = rank == self.rank and getattr(self, "_returning", False)
            maps = self.group._exchange(None if returning else {int(k): int(v) for k, v in self.owners().items()})
            if returning:
                self._owner = {int(k): int(v) for k, v in next(m for m in maps if m is not None).items()}
                self._returning = False
            if self.owners().get(rank, rank) != rank:
                self.handback(self.group._exchange, [rank])
        if rank != self.rank:
            self.segments.pop(rank, None)
        self.owners()[rank] = rank
        self.set_members(sorted(set(self.members()) | {rank}))

    def sync_epoch(self) -> None:
        """Agree on the flag-token epoch across the group (control-plane all-gather).  Needed when a
        rank that missed launches (re)joins: stale flag words must never satisfy a new barrier."""
        vals = self.group._exchange(int(self.attempt))
        self.attempt = max(int(v) for v in vals)

    def admit(self, exchange, order: List[int], new: List[int], joining: bool) -> None:
        """Round boundary at which replacement PROCESSES join (collective over all ranks in ``order``, the
        survivors right after their outer step).  ``exchange`` is a control-plane all-gather that includes
        the newcomers (``peer_memory.store_exchange``; a torch process group cannot admit a process).
        Peer-memory mappings are re-bound, survivors hand the newcomers' shards back, round and flag epoch are
        agreed; a newcomer (``joining``) then pulls the current weights from the lowest surviving rank over
        NVLink and rebuilds master / anchor / its outer shard.  The last exchange keeps everybody from
        launching the next outer step before the newcomers are seeded."""
        new = sorted(int(r) for r in new)
        survivors = [r for r in order if r not in new]
        self.group.rebind(exchange, new, members=order)
        self.round = max(int(r) for r in exchange(int(self.round)))
        maps = exchange(None if joining else dict(self.owners()))       # who runs which dead rank's shard
        if joining:
            self._owner = {int(k): int(v) for k, v in next(m for m in maps if m is not None).items()}
            self.reseed_from(survivors[0])      # parameters over NVLink; theta_outer / momentum placeholders
            self._returning = False             # the owner map came with `maps` above
            torch.cuda.synchronize()
        self.handback(exchange, new)            # the holders' fp32 theta_outer + momentum go to the newcomers
        if joining:
            self._owner[self.rank] = self.rank
            self.set_members(order)
        else:
            for r in new:
                self.readmit(r, handback=False)
        self.sync_epoch()
        exchange("ready")

    def reseed_from(self, src_rank: int) -> None:
        """(Re)join: pull the current global weights from a live rank over peer memory and rebuild the local
        state that derives from them.  theta_outer / momentum of the own shard are only
        placeholders here (theta from the pulled weights, momentum zero): `readmit` / `admit` replace them with
        the holder's exact fp32 state (`handback`) whenever somebody kept the shard going.  The inner AdamW
        moments are left alone: a returning rank keeps its own, a new process starts from zero anyway."""
        self.group.pull(src_rank, self.param_buf.name)
        C = load_C()
        C.bf16_to_master(self.opt.params.data_ptr(), self.opt.master.data_ptr(), self.n, cuda_stream())
        self.anchor_buf.local.copy_(self.opt.params)
        sl = slice(self.shard_begin, self.shard_begin + self.shard_len)
        self.theta_outer.copy_(self.opt.master[sl])
        self.momentum.zero_()
        self._returning = True          # `readmit` fetches the owner map (and the shard's state) from the others

    @staticmethod
    def merge_verdicts(verdicts: List[tuple]) -> tuple:
        """One verdict for all survivors from their per-rank ``(status, missing mask)`` of a fused step.

        The barrier time-outs are per rank, so the ranks of one launch can disagree: a slow rank may arrive inside one
        survivor's window and after another's.  A rank that PASSED barrier 1 has run phase B -- it changed its
        theta_outer / momentum and stored new parameters into its peers' buffers -- so a mix of "completed" (0) or
        "failed mid-broadcast" (2) with "timed out at barrier 1" (1) cannot be retried by dropping ranks: every rank
        must abort and restore.  Only when NOBODY got past barrier 1 is the state untouched everywhere; the missing
        ranks are then the union of what the survivors saw."""
        if not verdicts:
            return (0, 0)
        sts = [int(v[0]) for v in verdicts]
        missing = 0
        for v in verdicts:
            missing |= int(v[1])
        if all(s == 0 for s in sts):
            return (0, 0)
        if all(s == 1 for s in sts):
            return (1, missing)
        return (2, missing)

    def step_resilient(self, max_retries: int = 2, agree=None) -> tuple:
        """``step()`` that survives ranks that never show up: on a barrier-1 timeout the missing
        ranks are dropped, their shards adopted, and the step is retried with the survivors.

        ``agree``: a control-plane all-gather among the ranks that are still alive (``agree(obj) -> [obj of every
        rank that answered]``, e.g. the executor's membership exchange).  With it the survivors settle on ONE verdict
        (`merge_verdicts`) before anyone retries, so that a rank which saw the step fail mid-broadcast stops all of
        them; without it each rank acts on its own status (single-process drills)."""
        for _ in range(max_retries + 1):
            rnd = self.round
            self.step()
            st, missing = self.check()
            if agree is not None:
                st, missing = self.merge_verdicts([tuple(v) for v in agree((int(st), int(missing)))])
                self.last_status = (st, missing)
            if st == 0:
                return (0, 0)
            self.round = rnd                       # the failed attempt did not advance the round
            if st != 1:
                raise RuntimeError(f"outer step failed mid-broadcast (status {st}, missing mask {missing:#x}); "
                                   f"restore from the last checkpoint")
            dead = [r for r in range(self.world) if (missing >> r) & 1]
            self.drop_ranks(dead)
        raise RuntimeError("outer step kept timing out")

    def _pipeline_shape(self) -> tuple:
        """(num_chunks, link_ctas), tuned on 8xB200 (profiles/outer_sweep_n8_v2.json).  Pipelining
        pays once a chunk's HBM/NVLink time dwarfs the extra cross-GPU barriers (~15 us each): one
        chunk per 128 MB of delta, at most 16.  At N=2 phase B is itself HBM-bound, so splitting the
        SMs by role does not help and the plain A|B|C schedule is kept."""
        import os

        k = int(os.environ.get("HYPHA_OUTER_CHUNKS", self.cfg.pipeline_chunks))
        if k <= 0:
            k = 1 if self.world <= 2 else max(1, min(16, (2 * self.n) // (128 << 20)))
        b = int(os.environ.get("HYPHA_OUTER_LINK_CTAS", self.cfg.link_ctas))
        if b <= 0:
            # CTAs that keep NVLink busy while the rest stream HBM: the multimem path needs fewer
            # requests in flight (one load returns the reduced value) than 8 P2P loads per vector
            b = 66 if self.world <= 2 else (32 if self._multicast_ptrs()[0] else 48)
        return (k, b)

    def _multicast_ptrs(self) -> tuple:
        """(mc_delta, mc_param) when both buffers are NVLS-bound and every rank participates; the
        switch cannot skip a member, so a shrunk membership falls back to P2P loads/stores."""
        full = self.member_mask == (1 << self.world) - 1 and not getattr(getattr(self, "group", None), "multicast_broken", False)
        md = getattr(self.delta_buf, "mc_ptr", 0)
        mp = getattr(self.param_buf, "mc_ptr", 0)
        if full and md and mp and self.use_nvls:
            return (md, mp)
        return (0, 0)

    def check(self) -> tuple:
        """Host-side status of the last fused step: (status, missing_rank_mask)."""
        if self.backend != "fused":
            return (0, 0)
        torch.cuda.current_stream().synchronize()
        self.last_status = tuple(load_C().diloco_outer_step_status())
        return self.last_status

    def wire_bytes_per_step(self) -> int:
        """bf16 bytes this rank injects + receives over NVLink per outer step (fused path)."""
        if self.world == 1:
            return 0
        return 2 * 2 * self.shard_len * (self.world - 1)

    def state_dict(self) -> Dict[str, torch.Tensor]:
        sd = {"theta_outer": self.theta_outer, "momentum": self.momentum,
              "round": torch.tensor(self.round), "shard_begin": torch.tensor(self.shard_begin),
              "world": torch.tensor(self.world)}
        if self.backend == "fused":
            # the pseudo-gradient is theta - anchor, anchor = the weights right after the last outer step (bf16,
            # full length, one copy per rank): part of the state unless the checkpoint is taken exactly at a
            # round boundary.  Shards adopted from dropped ranks are NOT saved: checkpoint with full membership.
            sd["anchor"] = self.anchor_buf.local
        elif self.backend == "sharded":
            sd["anchor"] = self.anchor
        return sd

    def load_state_dict(self, sd) -> None:
        # a shard only makes sense at the place and team size it was saved for
        if "world" in sd and int(sd["world"]) != self.world:
            raise ValueError(f"outer state was saved by a team of {int(sd['world'])}, this one has {self.world} ranks")
        if "shard_begin" in sd and int(sd["shard_begin"]) != self.shard_begin:
            raise ValueError(f"outer state belongs to the shard at {int(sd['shard_begin'])}, this rank owns "
                             f"{self.shard_begin} (rank order changed?)")
        if tuple(sd["theta_outer"].shape) != tuple(self.theta_outer.shape):
            raise ValueError(f"outer shard has {tuple(sd['theta_outer'].shape)} elements, expected {tuple(self.theta_outer.shape)}")
        if self.backend == "fused" and len(getattr(self, "segments", {})) > 1:
            raise ValueError("cannot load outer state while this rank holds adopted shards: restore with the full membership")
        self.theta_outer.copy_(sd["theta_outer"])
        self.momentum.copy_(sd["momentum"])
        self.round = int(sd["round"])
        if self.backend == "fused":
            if "anchor" in sd:
                self.anchor_buf.local.copy_(sd["anchor"])
            else:       # checkpoint written before the anchor was saved: only right at a round boundary
                self.anchor_buf.local.copy_(self.opt.params)
        elif self.backend == "sharded":
            self.anchor.copy_(sd["anchor"] if "anchor" in sd else self.opt.params)
