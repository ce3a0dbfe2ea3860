"""DiLoCo training executor: ``python -m hypha_b200.executor.training --socket S --work-dir W --job J``.

Same contract as the reference executor (executors/accelerate/.../training.py:28-162): talk to
the worker's job bridge over a Unix socket, fetch model + data, run inner AdamW steps while
reporting ``{"status": {"batch_size": n}}`` after every batch, and when the scheduler answers
``ScheduleUpdate{counter}`` count down, ship the pseudo-gradient ``theta - theta_prev`` to the
aggregator, wait for the outer update, apply ``theta = theta_prev + update`` and acknowledge.

What is different:
  * theta_prev is kept in memory (device) instead of a safetensors file that is re-read twice per
    round, and the model never leaves the device (Q11);
  * the loss is accumulated on the device and read back once per round (Q10);
  * ``round_time`` is dropped (Q2); models are local architectures (Q12); WSD reads
    ``decay_steps`` (Q4);
  * on CUDA the Llama family trains with the sm_100a kernels + fused flat AdamW.
"""
from __future__ import annotations

import argparse
import json
import os
import time
from typing import Any, Dict, Iterator, List, Optional, Tuple

import torch

from .api import Session
from .dataset import IterableStreamDataSet, forever
from .model import _LlamaAdapter, get_model, get_preprocessor

FETCH_PATH = "artifacts"


def prepare_files(config: dict, session: Session) -> None:
    session.fetch(config["model"]["artifact"])
    if config.get("preprocessor"):
        session.fetch(config["preprocessor"]["artifact"])


def fetch_data(session: Session, data: Any, work_dir: str) -> Iterator[str]:
    while True:
        got = session.fetch(data)
        yield os.path.join(work_dir, got[0]["path"])


def accumulation_of(config: dict) -> Tuple[int, str]:
    """(micro-batches per inner step, main-gradient type) of a train executor's job config
    (`[scheduler.job.precision] micro_batches / main_grad`; absent: one micro-batch, the reference's behaviour)."""
    a = config.get("accumulation") or {}
    k, main_grad = int(a.get("micro_batches", 1)), str(a.get("main_grad", "fp32"))
    if k < 1 or main_grad not in ("fp32", "bf16"):
        raise ValueError(f"bad accumulation spec {a!r}")
    return k, main_grad


def split_batch(batch: Dict[str, torch.Tensor], k: int) -> List[Dict[str, torch.Tensor]]:
    """``k`` equal micro-batches along the sample dimension.  A batch whose size is not a multiple of ``k`` (the
    tail of a data slice) is cut into the largest number of EQUAL parts <= k instead, so that the mean over
    micro-batches stays the mean over samples."""
    n = int(next(iter(batch.values())).shape[0])
    k = max(1, min(k, n))
    while n % k:
        k -= 1
    if k == 1:
        return [batch]
    return [{key: v[i * (n // k):(i + 1) * (n // k)] for key, v in batch.items()} for i in range(k)]


class InnerLoop:
    """Model + inner optimizer + the bookkeeping DiLoCo needs (snapshot / delta / merge)."""

    def __init__(self, model: torch.nn.Module, opt_cfg: dict, sched_cfg: Optional[dict], device: str,
                 accumulation: Tuple[int, str] = (1, "fp32")):
        from ..parallel.diloco import AdamWConfig, FlatAdamW, make_lr_lambda

        self.model, self.device = model, device
        self.micro_batches, self.main_grad_type = accumulation
        self._main_grad = None
        self.flat = isinstance(model, _LlamaAdapter) and device.startswith("cuda")
        lr = float(opt_cfg["learning-rate"])
        betas = tuple(opt_cfg["betas"]) if opt_cfg.get("betas") else (0.9, 0.999)
        eps = float(opt_cfg["epsilon"]) if opt_cfg.get("epsilon") else 1e-8
        lam = make_lr_lambda(sched_cfg)
        if self.flat:
            net = model.net
            self.opt = FlatAdamW(net.flat.data, net.flat.grad, AdamWConfig(lr=lr, betas=betas, eps=eps,
                                                                          state_dtype="fp32"), lam)
            self.prev = self.opt.master.clone()
        else:
            self.opt = torch.optim.AdamW(model.parameters(), lr=lr, betas=betas, eps=eps)
            self.sched = torch.optim.lr_scheduler.LambdaLR(self.opt, lam)
            self.prev = {k: v.detach().clone() for k, v in model.state_dict().items()}

    def _loss(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        extra = {}
        if self.flat and batch.get("labels") is not None and not batch["labels"].is_cuda:
            # shift the labels while they are still on the host: the count of scored positions comes for free there
            from .model import hf_shift_labels

            targets, extra["n_valid"] = hf_shift_labels(batch["labels"])
            batch = dict({k: v for k, v in batch.items() if k != "labels"}, targets=targets)
        batch = {k: v.to(self.device, non_blocking=True) for k, v in batch.items()}
        out = self.model(**batch, **extra)
        return out if isinstance(out, torch.Tensor) else out["loss"]

    def step(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        micro = split_batch(batch, self.micro_batches) if self.micro_batches > 1 else [batch]
        if len(micro) > 1:
            return self._step_accumulated(micro)
        if not self.flat:
            self.opt.zero_grad(set_to_none=True)
        loss = self._loss(batch)
        loss.backward()
        self.opt.step()
        if not self.flat:
            self.sched.step()
        return loss.detach()

    def _step_accumulated(self, micro: List[Dict[str, torch.Tensor]]) -> torch.Tensor:
        """One optimizer step on the mean gradient of the micro-batches (`[scheduler.job.precision]
        micro_batches`).  torch models: autograd sums ``loss / k`` into the fp32 ``.grad``.  Flat sm_100a model:
        the bf16 gradient buffer is summed in fp32 (`parallel.main_grad.MainGrad`) or, with ``main_grad =
        "bf16"``, by the accumulate mode of the wgrad GEMM epilogues with 1/k folded into the AdamW launch."""
        k = len(micro)
        total = None
        if not self.flat:
            self.opt.zero_grad(set_to_none=True)
            for mb in micro:
                loss = self._loss(mb)
                (loss / k).backward()
                total = loss.detach() if total is None else total + loss.detach()
            self.opt.step()
            self.sched.step()
            return total / k
        net, fp32 = self.model.net, self.main_grad_type == "fp32"
        if fp32 and self._main_grad is None:
            from ..parallel.main_grad import MainGrad

            self._main_grad = MainGrad(net.flat.grad)
        try:
            for i, mb in enumerate(micro):
                net.set_grad_accumulate(not fp32 and i > 0)
                loss = self._loss(mb)
                loss.backward()
                if fp32:
                    self._main_grad.add()
                total = loss.detach() if total is None else total + loss.detach()
        finally:
            net.set_grad_accumulate(False)
        if fp32:
            self._main_grad.finalize()
            self.opt.step()
        else:
            self.opt.step(grad_scale=1.0 / k)
        return total / k

    def delta_state_dict(self) -> Dict[str, torch.Tensor]:
        """theta - theta_prev per tensor, fp32 on CPU (utils.py:118-123)."""
        if self.flat:
            d = (self.opt.master - self.prev)
            net = self.model.net
            return {n: net.flat.view(n, of=d).detach().float().cpu().contiguous() for n in net.flat.layout}
        sd = self.model.state_dict()
        return {k: (v.detach().float() - self.prev[k].float()).cpu().contiguous() for k, v in sd.items()
                if v.dtype.is_floating_point}

    def apply_update(self, update: Dict[str, torch.Tensor]) -> None:
        """theta = theta_prev + update; becomes the new theta_prev (utils.py:105-115)."""
        if self.flat:
            net = self.model.net
            for n in net.flat.layout:
                net.flat.view(n, of=self.prev).add_(update[n].to(self.prev.device).view(net.flat.layout[n][1]))
            self.opt.master.copy_(self.prev)
            net.flat.data.copy_(self.prev)
            return
        sd = self.model.state_dict()
        with torch.no_grad():
            for k, v in sd.items():
                if k in update:
                    new = self.prev[k].float() + update[k].to(v.device)
                    v.copy_(new.to(v.dtype))
                    self.prev[k] = v.detach().clone()


def _step_delay() -> float:
    """Fault-injection / test knob: HYPHA_EXECUTOR_STEP_DELAY_MS makes every inner step that much slower
    (a "slow worker" for the performance-aware scheduling tests); 0 in production."""
    try:
        return max(0.0, float(os.environ.get("HYPHA_EXECUTOR_STEP_DELAY_MS", "0"))) / 1e3
    except ValueError:
        return 0.0


class StatusPipe:
    """Per-batch `status{batch_size}` reports without stalling the training loop (the
    reference does a synchronous HTTP -> RPC round trip after every batch).  The report for batch k travels
    on a helper thread while batch k+1 computes; its answer is read one batch later.  A
    `ScheduleUpdate{counter}` answer to batch k-1 means "counter more batches after k-1": batch k was the
    first of them, so counter-1 remain -- the worker still runs exactly the number of batches the
    scheduler projected.  ``HYPHA_SYNC_STATUS=1`` restores the blocking behaviour.

    ``counter``: < 0 unknown, > 0 batches still to run, 0 synchronise now."""

    def __init__(self, session: Session):
        from concurrent.futures import ThreadPoolExecutor

        self.session = session
        self.sync = os.environ.get("HYPHA_SYNC_STATUS", "0") == "1"
        self._pool = None if self.sync else ThreadPoolExecutor(max_workers=1, thread_name_prefix="status")
        self._pending = None

    @staticmethod
    def _checked(resp: dict) -> dict:
        if resp["type"] == "Error":
            raise RuntimeError("scheduler rejected status")
        return resp

    def _harvest(self) -> Optional[dict]:
        if self._pending is None:
            return None
        resp, self._pending = self._checked(self._pending.result()), None
        return resp

    def report(self, batch_size: int, counter: int) -> int:
        """Report one finished batch; returns the updated countdown."""
        msg = {"status": {"batch_size": batch_size}}
        if self.sync:
            resp = self._checked(self.session.send_status(msg))
            return int(resp["counter"]) if resp["type"] == "ScheduleUpdate" else counter - 1
        prev = self._harvest()
        if counter > 0:
            counter -= 1                                   # this batch counts against the known countdown
        elif prev is not None and prev["type"] == "ScheduleUpdate":
            counter = max(0, int(prev["counter"]) - 1)     # the batch just finished was the first of them
        if counter == 0:
            self._checked(self.session.send_status(msg))   # last batch: report it before `update`
        else:
            self._pending = self._pool.submit(self.session.send_status, msg)
        return counter

    def drain(self) -> None:
        self._harvest()


# ---------------------------------------------------------------------------------------------
# collective mode: the train workers are the ranks of one process group and run the outer step
# among themselves -- no aggregate worker, no files (TrainExecutorConfig.collective)
# ---------------------------------------------------------------------------------------------
class CollectiveBroken(RuntimeError):
    """A rank died while the pseudo-gradients were being all-reduced; nothing was applied yet."""


class _Group:
    """The ranks an outer step is reduced over.  On GPUs this is the default torch.distributed (NCCL)
    group.  On CPU it is a gloo group this object OWNS (its own TCPStore + ProcessGroupGloo, no global
    state): it can be abandoned and re-created with other members while the job runs, and its collectives
    carry a short timeout so that a rank blocked on a peer that left fails fast instead of hanging."""

    def __init__(self, kind: str, rendezvous: Optional[str], rank: int, world: int):
        self.kind, self.rank, self.world = (kind if world > 1 else "solo"), rank, world
        self.pg = self.store = None
        if self.kind == "gloo":
            import torch.distributed as dist
            from datetime import timedelta

            host, port = rendezvous[len("tcp://"):].rsplit(":", 1)
            op_timeout = float(os.environ.get("HYPHA_COLLECTIVE_TIMEOUT_S", "20"))
            self.store = dist.TCPStore(host, int(port), world, rank == 0, timedelta(seconds=120))
            self.pg = dist.ProcessGroupGloo(self.store, rank, world, timedelta(seconds=op_timeout))

    def all_reduce(self, t: torch.Tensor, op: str = "sum") -> None:
        if self.kind == "solo":
            return
        import torch.distributed as dist

        rop = dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM
        try:
            if self.kind == "gloo":
                opts = dist.AllreduceOptions()
                opts.reduceOp = rop
                self.pg.allreduce([t], opts).wait()
            else:
                dist.all_reduce(t, op=rop)
        except RuntimeError as exc:
            raise CollectiveBroken(str(exc)) from exc

    def broadcast(self, t: torch.Tensor, root: int = 0) -> None:
        if self.kind == "solo":
            return
        import torch.distributed as dist

        try:
            if self.kind == "gloo":
                opts = dist.BroadcastOptions()
                opts.rootRank = root
                self.pg.broadcast([t], opts).wait()
            else:
                dist.broadcast(t, root)
        except RuntimeError as exc:
            raise CollectiveBroken(str(exc)) from exc

    def abandon(self) -> None:
        """Drop an owned group without blocking: the C++ destructors (which close the sockets and thereby
        fail the peers' pending operations) run on a helper thread."""
        if self.kind != "gloo":
            return
        import threading

        refs = [self.pg, self.store]
        self.pg = self.store = None
        self.kind = "abandoned"
        threading.Thread(target=refs.clear, daemon=True).start()


class ReplicatedNesterov:
    """Outer step for the CPU / generic-model plumbing path: all-reduce the pseudo-gradient over the
    job's ranks, true 1/N mean, Nesterov with the reference's numerics (parameter_server.rs:386-446);
    every rank keeps the full outer state.

    Two phases so that a rank dying mid-collective cannot leave the model half-updated: first every
    tensor's mean is reduced into a temporary (on failure nothing has changed and the step can be retried
    with the surviving ranks), then momentum / parameters are updated locally."""

    def __init__(self, lr: float, mu: float):
        self.lr, self.mu = lr, mu
        self.m: Optional[Dict[str, torch.Tensor]] = None

    @staticmethod
    def _mean(g: torch.Tensor, group: _Group, exclude: int = 0, weight: Optional[float] = None) -> torch.Tensor:
        """``weight`` = this rank's share of the round's samples (already normalised over the group): the mean
        over samples is then just the sum of the weighted pseudo-gradients."""
        if weight is not None:
            g.mul_(weight)
            if group.world > 1:
                group.all_reduce(g)
            return g
        if group.world > 1:
            group.all_reduce(g)
            g.div_(max(1, group.world - exclude))      # ranks that just joined contribute a zero delta
        return g

    @staticmethod
    def sample_share(samples: int, group: _Group, device: str = "cpu") -> Optional[float]:
        """samples / (sum over the group); None when nobody processed anything (all ranks just joined)."""
        t = torch.tensor([float(samples)], device=device)
        if group.world > 1:
            group.all_reduce(t)
        total = float(t.item())
        return (float(samples) / total) if total > 0 else None

    @staticmethod
    def _check_round(round_no: int, group: _Group, device: str = "cpu") -> None:
        """All ranks must be reducing the same round (guards the corner where one survivor finished an
        interrupted collective and moved on while another one retries): fail loudly, never mix rounds."""
        if group.world <= 1 or round_no < 0:
            return
        tag = torch.tensor([float(round_no), float(-round_no)], device=device)     # NCCL needs device tensors
        group.all_reduce(tag, "max")
        if int(tag[0]) != round_no or int(-tag[1]) != round_no:
            raise RuntimeError(f"ranks disagree on the round being reduced ({int(-tag[1])}..{int(tag[0])} vs {round_no})")

    @torch.no_grad()
    def step(self, loop: "InnerLoop", group, round_no: int = -1, exclude: int = 0,
             samples: Optional[int] = None) -> None:
        """``samples``: this rank's sample count of the round -> mean over samples (`aggregation.weighting =
        "samples"`); None -> mean over ranks."""
        if isinstance(group, int):                       # plain world size: single rank / tests
            group = _Group("solo", None, 0, 1) if group <= 1 else _Group("default", None, 0, group)
        first = self.m is None
        self._check_round(round_no, group, getattr(loop, "device", "cpu"))
        weight = None if samples is None else self.sample_share(samples, group, getattr(loop, "device", "cpu"))
        if weight is not None:
            exclude = 0                                  # newcomers have zero samples: their weight is zero already
        if loop.flat:
            net = loop.model.net
            g = self._mean(loop.opt.master - loop.prev, group, exclude, weight)
            m = g.clone() if first else self.m["flat"].mul_(self.mu).add_(g)
            self.m = {"flat": m}
            loop.prev.add_(self.lr * (self.mu * m + g))
            loop.opt.master.copy_(loop.prev)
            net.flat.data.copy_(loop.prev)
            return
        sd = loop.model.state_dict()
        means = {k: self._mean(v.detach().float() - loop.prev[k].float(), group, exclude, weight)  # phase 1: may raise
                 for k, v in sd.items() if v.dtype.is_floating_point}
        if first:
            self.m = {}
        for k, g in means.items():                                                       # phase 2: local
            m = g.clone() if first else self.m[k].mul_(self.mu).add_(g)
            self.m[k] = m
            v = sd[k]
            v.copy_((loop.prev[k].float() + self.lr * (self.mu * m + g)).to(v.dtype))
            loop.prev[k] = v.detach().clone()


def gpu_join_enabled() -> bool:
    """Replacement processes may join the fused GPU engine (`_TrainerEngine._admit`): on by default since the
    path ran on hardware (tests/mgpu_join_check.py and tests/mgpu_box_fault_check.py: `kill -9` of an executor,
    replacement leased, admitted over CUDA IPC / VMM descriptors, weights pulled over NVLink);
    ``HYPHA_GPU_JOIN=0`` turns it off (the job then continues with the survivors only)."""
    return os.environ.get("HYPHA_GPU_JOIN", "1") == "1"


class _TrainerEngine:
    """Llama on a B200: the public DiLoCoTrainer (flat buffers in peer memory, fused AdamW, fused
    NVLink outer step)."""

    def __init__(self, local: str, config: dict, coll: dict, dev_index: int, joining: bool = False):
        from ..models.llama import LlamaConfig
        from ..parallel.diloco import AdamWConfig, DiLoCoConfig
        from ..parallel.peer_memory import PeerGroup
        from ..trainer import DiLoCoTrainer

        meta = json.load(open(os.path.join(local, "config.json")))
        cfg = LlamaConfig.from_hf_dict(meta)
        o = config["optimizer"]
        adamw = AdamWConfig(lr=float(o["learning-rate"]), betas=tuple(o["betas"]) if o.get("betas") else (0.9, 0.999),
                            eps=float(o["epsilon"]) if o.get("epsilon") else 1e-8)
        outer = coll.get("outer") or {}
        diloco = DiLoCoConfig(inner_steps=1 << 62, outer_lr=float(outer.get("learning-rate", 0.7)),
                              outer_momentum=float(outer.get("momentum", 0.9)))
        self.joining = joining          # a replacement PROCESS: maps its peers in _admit(), trains afterwards
        self._pending: List[int] = []   # slots of newcomers to admit after the coming outer step
        self._mem: Optional[dict] = None
        self._me = coll.get("peer")
        if joining:
            # the dead rank's number and the ORIGINAL world size: shard ownership is by rank
            group = PeerGroup.joining(int(coll["slot"]), int(coll["slots_world"]), dev_index)
        else:
            group = PeerGroup.from_torch_distributed(dev_index) if coll["world"] > 1 else None
        self.trainer = DiLoCoTrainer.create(cfg, adamw=adamw, diloco=diloco, group=group, device=dev_index,
                                            seed=int(meta.get("seed", 0)), scheduler=config.get("scheduler"))
        from .model import load_weights

        sd = load_weights(local)                    # model.safetensors, or the shards of its index
        if sd is not None:
            self.trainer.model.load_hf_state_dict(sd)
            self.trainer.opt.master.copy_(self.trainer.model.flat.data)
            del sd
        self.backend = self.trainer.outer.backend
        self.micro_batches, self.main_grad_type = accumulation_of(config)

    def step(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        # the dataset follows the Hugging Face convention (labels aligned with input_ids, the shift happens in the
        # model): same targets as the CPU path's `_LlamaAdapter` and as the reference's `model(**batch)`
        from .model import hf_shift_labels

        batch = {"input_ids": batch["input_ids"], "labels": batch.get("labels", batch["input_ids"])}
        micro = split_batch(batch, self.micro_batches) if self.micro_batches > 1 else [batch]
        if len(micro) > 1:
            return self.trainer.inner_step_accumulated(
                [(mb["input_ids"],) + hf_shift_labels(mb["labels"]) for mb in micro],
                main_grad=self.main_grad_type, ignore_index=-100)
        ids = batch["input_ids"]
        targets, n_valid = hf_shift_labels(batch["labels"])
        return self.trainer.inner_step(ids, targets, ignore_index=-100, n_valid=n_valid)

    def regroup(self, membership: dict, coll: dict, **_) -> None:
        """The scheduler dropped ranks: shrink the fused kernel's membership and adopt orphaned shards now
        (`DiLoCoOuter.drop_ranks`) instead of discovering the dead rank by a barrier timeout.  Peer-memory
        mappings and the original rank numbering stay as they are."""
        outer = self.trainer.outer
        if outer.backend != "fused":
            return
        alive = set(membership["members"])
        slots = membership.get("slots") or {}
        self._mem = membership
        if self.joining:
            return                       # nothing mapped yet: everything happens in _admit()
        # ranks by slot (stable numbering); the dispatch-time peer list gives the slots of the first generation
        by_slot = {i: p for i, p in enumerate(coll.get("peers", []))}
        by_slot.update({int(v): k for k, v in slots.items()})
        joined = [int(slots[p]) for p in (membership.get("joined") or []) if p in slots]
        dead = [i for i in outer.members() if by_slot.get(i) not in alive or i in joined]
        if dead:
            outer.drop_ranks(dead)
            print(f"membership generation {membership['generation']}: continuing without ranks {dead}", flush=True)
        if joined and gpu_join_enabled():
            self._pending = joined

    def _admit(self) -> None:
        """Round boundary with newcomers (HYPHA_GPU_JOIN=1): the survivors have just finished their outer step.
        All members of the new generation meet on a TCPStore at the generation's rendezvous (no process group:
        a torch group cannot admit a process), re-bind the peer-memory mappings (`PeerGroup.rebind`), hand the
        newcomers' shards back and align the flag epoch; a newcomer then pulls the current weights from the
        lowest surviving rank over NVLink and rebuilds master / anchor / its outer shard (`reseed_from`)."""
        from datetime import timedelta
        from urllib.parse import urlparse

        import torch.distributed as dist

        from ..parallel.peer_memory import store_exchange

        mem, outer = self._mem, self.trainer.outer
        slots = {k: int(v) for k, v in (mem.get("slots") or {}).items() if k in mem["members"]}
        order = sorted(slots.values())
        me = outer.rank
        new = sorted(int(slots[p]) for p in (mem.get("joined") or []) if p in slots)
        survivors = [r for r in order if r not in new]      # the lowest one hosts the store
        u = urlparse(mem["rendezvous"])
        store = dist.TCPStore(u.hostname, u.port, len(order), me == survivors[0], timedelta(seconds=120),
                              wait_for_workers=False)
        exchange = store_exchange(store, order.index(me), len(order), prefix=f"gen{mem['generation']}")
        outer.admit(exchange, order, new, self.joining)
        self._store = store               # keep the exchange alive for later sync_epoch calls
        print(f"membership generation {mem['generation']}: ranks {new} admitted into the fused engine", flush=True)
        self._pending, self.joining = [], False

    # -- job checkpoint through the scheduler (`[scheduler.job.checkpoint]`; the reference has none) --------
    @torch.no_grad()
    def save_checkpoint(self, ckpt_dir: str, round_no: int, work_dir: str) -> Optional[str]:
        """Right after an outer step every rank holds the same bf16 parameters, and rank r the fp32 theta_outer +
        momentum of shard r (ZeRO-1).  Rank 0 writes the parameters, every rank its outer shard; rank 0 announces
        the round once all pieces are on disk.  As in the other aggregation modes the inner AdamW state is not
        part of a job checkpoint (`DiLoCoTrainer.save_checkpoint` is the full-state, per-rank variant)."""
        from safetensors.torch import save_file

        from ..worker.jobs import publish_sharded_checkpoint

        outer = self.trainer.outer
        if outer.backend == "fused" and outer.members() != list(range(outer.world)):
            print(f"checkpoint of round {round_no} skipped: membership is not full ({outer.members()})", flush=True)
            return None
        torch.cuda.synchronize()
        world, rank = outer.world, outer.rank
        names = ["collective_params"] + [f"outer_shard_{r}" for r in range(world)]
        tmp = os.path.join(work_dir, f"outer_shard_{rank}.tmp")
        save_file({"theta_outer": outer.theta_outer.detach().cpu(), "momentum": outer.momentum.detach().cpu(),
                   "meta": torch.tensor([outer.round, outer.shard_begin, outer.shard_len, world], dtype=torch.int64)}, tmp)
        where = publish_sharded_checkpoint(ckpt_dir, round_no, f"outer_shard_{rank}", tmp,
                                           expect=None)
        os.unlink(tmp)
        if rank == 0:
            tmp = os.path.join(work_dir, "collective_params.tmp")
            save_file({"params": self.trainer.model.flat.data.detach().cpu()}, tmp)
            where = publish_sharded_checkpoint(ckpt_dir, round_no, "collective_params", tmp, expect=names)
            os.unlink(tmp)
        return where

    @torch.no_grad()
    def load_checkpoint(self, path: str) -> None:
        from safetensors.torch import load_file

        tr, outer = self.trainer, self.trainer.outer
        dev = str(tr.device)
        tr.model.flat.data.copy_(load_file(os.path.join(path, "collective_params"), device=dev)["params"])
        tr.opt.master.copy_(tr.model.flat.data)                    # fp32 master restarts from the bf16 weights
        sh = load_file(os.path.join(path, f"outer_shard_{outer.rank}"), device=dev)
        rnd, begin, length, world = (int(x) for x in sh["meta"].tolist())
        if (begin, length, world) != (outer.shard_begin, outer.shard_len, outer.world):
            raise RuntimeError(f"checkpoint shard {outer.rank} covers [{begin}, +{length}) of a {world}-rank job; this job "
                               f"needs [{outer.shard_begin}, +{outer.shard_len}) of {outer.world}")
        outer.theta_outer.copy_(sh["theta_outer"])
        outer.momentum.copy_(sh["momentum"])
        outer.round = rnd
        if outer.backend == "fused":
            outer.anchor_buf.local.copy_(tr.model.flat.data)
        elif outer.backend == "sharded":
            outer.anchor.copy_(tr.model.flat.data)
        from ..ops.functional import bump_fp8_weight_version

        bump_fp8_weight_version()

    def outer_step(self, round_no: int = -1, samples: Optional[int] = None) -> None:
        # `samples` is ignored: the fused kernel takes the mean over ranks (equal shares on a homogeneous box)
        outer = self.trainer.outer
        if self.joining:
            self._admit()                 # the survivors run this round's outer step without the newcomer
            return
        if outer.backend == "fused":
            # a rank that never shows up costs one barrier timeout, then the survivors adopt its shard
            # and go on (DiLoCoOuter.step_resilient); raises only if the step cannot be completed
            before = outer.members()
            outer.step_resilient()
            if outer.members() != before:
                print(f"outer step continued without ranks {sorted(set(before) - set(outer.members()))}", flush=True)
            if self._pending:
                self._admit()
        else:
            self.trainer.outer_step()


class _GenericEngine:
    def __init__(self, local: str, config: dict, coll: dict, device: str, group: _Group):
        model = get_model(local, config["model"]["task"], device)
        self.loop = InnerLoop(model, config["optimizer"], config.get("scheduler"), device, accumulation_of(config))
        outer = coll.get("outer") or {}
        self.outer = ReplicatedNesterov(float(outer.get("learning-rate", 0.7)), float(outer.get("momentum", 0.9)))
        self.group = group
        self.backend = "all-reduce"
        self._exclude = 0             # newcomers in the coming outer step (they contribute a zero delta)

    def step(self, batch):
        return self.loop.step(batch)

    @torch.no_grad()
    def save_checkpoint(self, ckpt_dir: str, round_no: int, work_dir: str) -> Optional[str]:
        """Collective-mode job checkpoint: every rank holds the same global model theta_prev and outer momentum
        after an outer step, so rank 0 of the current group writes them (`<dir>/round-<n>/collective_state`).
        The inner AdamW state is not part of it (as in parameter-server mode a resumed job starts its inner
        optimizers afresh)."""
        if self.loop.flat or self.group.rank != 0:
            return None
        from safetensors.torch import save_file

        from ..worker.jobs import publish_checkpoint

        sd = self.loop.model.state_dict()
        out = {f"prev/{k}": v.detach().cpu().contiguous() for k, v in self.loop.prev.items()}
        out.update({f"buf/{k}": v.detach().cpu().contiguous() for k, v in sd.items() if k not in self.loop.prev})
        out.update({f"m/{k}": v.detach().cpu().contiguous() for k, v in (self.outer.m or {}).items()})
        tmp = os.path.join(work_dir, "collective_state.tmp")
        save_file(out, tmp)
        dst = publish_checkpoint(ckpt_dir, round_no, {"collective_state": tmp})
        os.unlink(tmp)
        return dst

    @torch.no_grad()
    def load_checkpoint(self, path: str) -> None:
        from safetensors.torch import load_file

        st = load_file(os.path.join(path, "collective_state"))
        sd = self.loop.model.state_dict()
        m = {}
        for key, v in st.items():
            kind, k = key.split("/", 1)
            if kind == "m":
                m[k] = v.to(self.loop.device).float()
            elif k in sd:
                sd[k].copy_(v.to(sd[k].dtype))
                if kind == "prev":
                    self.loop.prev[k] = sd[k].detach().clone()
        self.outer.m = m or None

    @torch.no_grad()
    def welcome(self, membership: dict, me: Optional[str] = None) -> None:
        """Right after a regroup that admitted newcomers: rank 0 (always a survivor -- newcomers are appended
        to the member list) broadcasts the current global model theta_prev and the outer momentum; every rank
        takes part, the newcomers adopt what they receive."""
        if self.loop.flat:
            raise RuntimeError("collective join is implemented for the replicated (non-flat) engine only")
        joined = list(membership.get("joined") or [])
        if not joined:
            return
        flag = torch.tensor([0.0 if self.outer.m is None else 1.0])
        self.group.broadcast(flag)
        has_m = bool(flag.item())
        sd = self.loop.model.state_dict()
        if has_m and self.outer.m is None:
            self.outer.m = {}
        for k in sorted(k for k, v in sd.items() if v.dtype.is_floating_point):
            prev = self.loop.prev[k].contiguous()
            self.group.broadcast(prev)
            self.loop.prev[k] = prev
            if me in joined:
                sd[k].copy_(prev)                  # newcomers start from the global model; survivors keep
                                                   # their locally trained weights (= their pseudo-gradient)
            if has_m:
                m = self.outer.m.get(k)
                if m is None:
                    m = torch.zeros_like(prev, dtype=torch.float32)
                m = m.contiguous()
                self.group.broadcast(m)
                self.outer.m[k] = m
        self._exclude = len(joined)
        print(f"welcomed {joined}: model and outer momentum broadcast from rank 0", flush=True)

    def abort_group(self) -> None:
        """Leave the group NOW (owned gloo groups only): closing our side makes the peers' pending
        collective fail too instead of leaving them blocked on us."""
        self.group.abandon()

    def regroup(self, membership: dict, coll: dict, use_cuda: bool = False, device: str = "cpu") -> None:
        """The scheduler changed the membership: form a new group with the surviving ranks on the fresh
        rendezvous address (CPU / gloo; the NCCL group of a GPU job is fixed)."""
        members = list(membership["members"])
        me = coll.get("peer")
        if me not in members:
            raise RuntimeError("this worker was removed from the job")
        if use_cuda:
            raise RuntimeError("membership changed: NCCL-backed generic jobs cannot regroup")
        self.group.abandon()
        self.group = _Group("gloo", membership["rendezvous"], members.index(me), len(members))
        print(f"membership generation {membership['generation']}: regrouped as rank {members.index(me)} of "
              f"{len(members)}", flush=True)

    def outer_step(self, round_no: int = -1, samples: Optional[int] = None) -> None:
        self.outer.step(self.loop, self.group, round_no, self._exclude, samples)
        self._exclude = 0


def main_collective(session: Session, config: dict, work_dir: str) -> None:
    import torch.distributed as dist

    coll = config["collective"]
    rank, world = int(coll["rank"]), int(coll["world"])
    use_cuda = torch.cuda.is_available() and os.environ.get("HYPHA_EXECUTOR_DEVICE", "cuda") != "cpu"
    dev_index = int(os.environ.get("HYPHA_DEVICE", rank % max(1, torch.cuda.device_count()))) if use_cuda else -1
    if use_cuda:
        torch.cuda.set_device(dev_index)
    device = f"cuda:{dev_index}" if use_cuda else "cpu"
    gpu_join = bool(coll.get("join")) and use_cuda
    if gpu_join and not gpu_join_enabled():
        raise RuntimeError("joining a running collective job on GPUs is switched off (HYPHA_GPU_JOIN=0)")
    if world > 1 and use_cuda and not gpu_join:
        from datetime import timedelta

        # GPUs: the default NCCL group (the peer-memory handle exchange of the fused path runs over it)
        timeout_s = float(os.environ.get("HYPHA_COLLECTIVE_TIMEOUT_S", "300"))
        dist.init_process_group("nccl", init_method=coll["rendezvous"], rank=rank, world_size=world,
                                timeout=timedelta(seconds=timeout_s), device_id=torch.device(device))
    prepare_files(config, session)
    local = os.path.join(work_dir, FETCH_PATH)
    family = json.load(open(os.path.join(local, "config.json"))).get("model_type", "")
    from ..models.llama import SM100_FAMILIES

    if use_cuda and family in SM100_FAMILIES:
        engine = _TrainerEngine(local, config, coll, dev_index, joining=gpu_join)
    elif coll.get("join"):
        engine = _GenericEngine(local, config, coll, device, _Group("solo", None, 0, 1))   # regrouped at its first barrier
    else:
        group = _Group("default" if use_cuda else "gloo", coll["rendezvous"], rank, world)
        engine = _GenericEngine(local, config, coll, device, group)
    print(f"collective rank {rank}/{world} on {device}: outer step via {engine.backend}", flush=True)
    pre_cfg = config.get("preprocessor")
    loader = torch.utils.data.DataLoader(
        IterableStreamDataSet(fetch_data(session, config["data"], work_dir), config["model"]["input-names"],
                              pre_cfg["input-names"] if pre_cfg else [], get_preprocessor(pre_cfg, local)),
        batch_size=config["batch_size"])
    batches = forever(loader)
    round_no = 1
    def sync():
        if use_cuda:
            torch.cuda.synchronize()

    status = StatusPipe(session)
    delay = _step_delay()
    generation = 0
    joining = bool(coll.get("join"))        # a replacement rank: no training before its first barrier
    ckpt = coll.get("checkpoint") or {}
    if ckpt.get("resume_round"):
        engine.load_checkpoint(os.path.join(ckpt["dir"], f"round-{int(ckpt['resume_round'])}"))
        round_no = int(ckpt["resume_round"]) + 1
        print(f"resumed from the collective checkpoint of round {round_no - 1}", flush=True)

    while True:
        loss_sum = torch.zeros((), device=device)
        steps, counter, tokens, samples = 0, (0 if joining else -1), 0, 0
        sync()
        t_round = time.perf_counter()
        while counter != 0:
            batch = next(batches)
            loss_sum += engine.step(batch)
            if delay:
                time.sleep(delay)
            steps += 1
            first = next(iter(batch.values()))
            bs = int(first.shape[0])
            samples += bs
            tokens += int(first.numel()) if first.dim() == 2 and not first.dtype.is_floating_point else bs
            counter = status.report(bs, counter)
        status.drain()
        sync()
        inner_s = time.perf_counter() - t_round
        # answered once every current member reached this point; tells us who takes part (a rank the
        # scheduler dropped in the meantime is no longer waited for)
        ack = session.send_status("update")
        mem = ack.get("membership") if isinstance(ack, dict) else None
        if mem and int(mem.get("generation", 0)) != generation:
            generation = int(mem["generation"])
            engine.regroup(mem, coll, use_cuda=use_cuda, device=device)
            if mem.get("joined") and hasattr(engine, "welcome"):
                engine.welcome(mem, coll.get("peer"))
        if joining:
            round_no = int(mem["round"]) + 1 if mem else round_no
            joining = False
            print(f"joined a running collective job at round {round_no}", flush=True)
        t_outer = time.perf_counter()
        for attempt in range(240):
            try:
                # every rank enters: the kernel / all-reduce is the barrier
                engine.outer_step(round_no, samples if config.get("weighting") == "samples" else None)
                break
            except CollectiveBroken as exc:
                # a rank died inside the collective: nothing was applied.  Ask the scheduler (it learns of the
                # death through the worker's JobStatus or the lost lease) until the membership changes, re-form
                # the group with the survivors and run the outer step again.
                print(f"outer step interrupted ({str(exc)[:80]}); waiting for the new membership", flush=True)
                if hasattr(engine, "abort_group"):
                    engine.abort_group()
                while True:
                    time.sleep(0.5)
                    ack = session.send_status("update")
                    mem = ack.get("membership") if isinstance(ack, dict) else None
                    if mem and int(mem.get("generation", 0)) != generation:
                        generation = int(mem["generation"])
                        engine.regroup(mem, coll, use_cuda=use_cuda, device=device)
                        break
                    attempt += 1
                    if attempt > 240:
                        raise
        sync()
        outer_ms = (time.perf_counter() - t_outer) * 1e3
        mean_loss = float(loss_sum.item() / max(1, steps))
        # the reference reports the loss only (training.py:148-150); the extra keys are SURVEY 5.5's
        # "tokens/s, step time, outer-step wall time" and travel through the same Metrics message
        if steps > 0:                              # a rank that joined at this barrier has nothing to report yet
            session.send_status({"metrics": {"round": round_no, "metrics": {
                "loss": mean_loss, "samples_or_tokens_per_s": tokens / max(inner_s, 1e-9),
                "inner_step_ms": inner_s * 1e3 / max(1, steps), "outer_step_ms": outer_ms,
                "inner_steps": float(steps), "tokens": float(tokens)}}})
        print(f"round {round_no}: {steps} inner steps, mean loss {mean_loss:.4f}, {inner_s * 1e3 / max(1, steps):.1f} ms/step; "
              f"outer step done in {outer_ms:.1f} ms", flush=True)
        # what the aggregate worker reports in parameter-server mode.  Every rank sends it (the scheduler
        # advances the round on the first one and ignores the rest), so no rank is a single point of
        # failure and each rank's own `updated` precedes its `update-received`.
        session.send_status("updated")
        if ckpt.get("dir") and round_no % max(1, int(ckpt.get("every", 1))) == 0 and hasattr(engine, "save_checkpoint"):
            where = engine.save_checkpoint(ckpt["dir"], round_no, work_dir)
            if where:
                print(f"collective checkpoint of round {round_no} written to {where}", flush=True)
        round_no += 1
        if session.send_status("update-received")["type"] == "Done":
            break
    print(f"Finished training of {round_no - 1} DiLoCo update rounds", flush=True)
    if world > 1 and use_cuda:
        # no closing barrier: nothing touches peer memory after the last outer step, and a barrier would
        # hang the survivors of a job that lost a rank
        try:
            dist.destroy_process_group()
        except Exception:
            pass


def apply_precision(precision: Optional[str], device: str) -> str:
    """Job-level GEMM precision (`[scheduler.job.precision] gemm`): ``mxfp8`` / ``fp8`` switch the Linear GEMMs of
    the sm_100a models to the block-scaled / per-tensor fp8 tensor-core kernels; on a CPU executor (and for models
    that do not run on the hand-written kernels) the setting has nothing to act on and bf16 / fp32 stays."""
    precision = precision or "bf16"
    if precision not in ("bf16", "mxfp8", "fp8"):
        raise ValueError(f"unknown GEMM precision {precision!r}")
    if precision != "bf16" and device.startswith("cuda"):
        from .. import ops

        ops.set_fp8(True, mode="mx" if precision == "mxfp8" else "tensor")
        print(f"inner-step GEMMs in {precision} ({'block-scaled tcgen05.mma.kind::mxf8f6f4' if precision == 'mxfp8' else 'kind::f8f6f4, per-tensor scales'})",
              flush=True)
        return precision
    return "bf16"


def main(socket_path: str, work_dir: str, job_json: str) -> None:
    from safetensors.torch import load_file, save_file

    if torch.cuda.is_available() and os.environ.get("HYPHA_EXECUTOR_DEVICE", "cuda") != "cpu":
        # one worker per GPU: the worker exports its `device` as HYPHA_DEVICE (config.WorkerConfig.device)
        index = int(os.environ.get("HYPHA_DEVICE", "0")) % max(1, torch.cuda.device_count())
        torch.cuda.set_device(index)
        device = f"cuda:{index}"
    else:
        device = "cpu"
    with Session(socket_path) as session:
        job = json.loads(job_json)
        executor = job["executor"]
        assert executor["class"] == "train"
        config = executor["config"]
        apply_precision(config.get("precision"), device)
        if config.get("collective"):
            return main_collective(session, config, work_dir)
        prepare_files(config, session)
        local = os.path.join(work_dir, FETCH_PATH)
        model = get_model(local, config["model"]["task"], device)
        loop = InnerLoop(model, config["optimizer"], config.get("scheduler"), device, accumulation_of(config))
        pre_cfg = config.get("preprocessor")
        loader = torch.utils.data.DataLoader(
            IterableStreamDataSet(fetch_data(session, config["data"], work_dir), config["model"]["input-names"],
                                  pre_cfg["input-names"] if pre_cfg else [], get_preprocessor(pre_cfg, local)),
            batch_size=config["batch_size"])
        batches = forever(loader)
        round_no = 1
        status = StatusPipe(session)
        delay = _step_delay()
        with session.receive(config["results"], "incoming") as receiver:
            updates = iter(receiver)
            if config.get("join"):
                # replacement worker: theta_0 is the seeded init every worker started from; the aggregate
                # worker sends U = sum of all outer updates so far, so theta_prev = theta_0 + U is the
                # current global model
                event = next(updates)
                latest = event[-1] if isinstance(event, list) else event
                rel = (latest.get("parameters") or {}).get("path") if isinstance(latest.get("parameters"), dict) else latest.get("path")
                cum = load_file(os.path.join(work_dir, rel))
                if cum:
                    loop.apply_update(cum)
                round_no = int(config["join"].get("after_round", 0)) + 1
                print(f"joined a running job after round {round_no - 1} ({len(cum)} tensors of cumulative update)", flush=True)
            while True:
                loss_sum = torch.zeros((), device=device)
                steps, counter, samples = 0, -1, 0
                while counter != 0:
                    batch = next(batches)
                    loss_sum += loop.step(batch)
                    if delay:
                        time.sleep(delay)
                    steps += 1
                    bs = int(next(iter(batch.values())).shape[0])
                    samples += bs
                    counter = status.report(bs, counter)
                status.drain()
                session.send_status("update")
                name = f"{round_no}_local_gradients.safetensors"
                delta = loop.delta_state_dict()
                if config.get("weighting") == "samples":
                    # the aggregate worker takes the mean over samples when every worker reports its count
                    delta["__samples__"] = torch.tensor([float(samples)], dtype=torch.float32)
                save_file(delta, os.path.join(work_dir, name))
                session.send_resource(config["updates"], name)
                session.send_status({"metrics": {"round": round_no, "metrics": {"loss": float(loss_sum.item() / max(1, steps))}}})
                print(f"round {round_no}: {steps} inner steps, mean loss {float(loss_sum.item() / max(1, steps)):.4f}; "
                      f"waiting for the outer update", flush=True)
                event = next(updates)
                latest = event[-1] if isinstance(event, list) else event
                rel = (latest.get("parameters") or {}).get("path") if isinstance(latest.get("parameters"), dict) else latest.get("path")
                loop.apply_update(load_file(os.path.join(work_dir, rel)))
                os.unlink(os.path.join(work_dir, name))
                round_no += 1
                if session.send_status("update-received")["type"] == "Done":
                    break
        print(f"Finished training of {round_no - 1} DiLoCo update rounds", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--socket", required=True)
    ap.add_argument("--work-dir", required=True)
    ap.add_argument("--job", required=True)
    a = ap.parse_args()
    main(a.socket, a.work_dir, a.job)
