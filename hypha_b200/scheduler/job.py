"""One DiLoCo job from the scheduler's point of view: allocate -> dispatch -> watch -> tear down.

Reference flow: crates/scheduler/src/bin/hypha-scheduler.rs:54-432.  What is added here: the
`collective` aggregation mode (no aggregate worker; the train workers' executors run the outer step
among themselves) and elasticity (a lost train worker is dropped, optionally replaced by a newcomer
that joins the running job) -- the reference exits on the first lost lease.
"""
from __future__ import annotations

import asyncio
import json
import os
import socket
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

from .. import messages as M
from ..transport import Network, NetworkError
from .allocator import GreedyWorkerAllocator, Task, TaskError, WorkerHandle
from .scheduling import (AimConnector, BatchSchedulerService, Connector as MetricsConnector, DataSchedulerService,
                         MetricsBridge, NoOpConnector)

ALLOCATION_PAUSE_S = 1.0        # hypha-scheduler.rs:243
REPLACE_ATTEMPTS = 4            # advertisements for a replacement worker before giving up
REPLACE_RETRY_S = 3.0


class JobError(Exception):
    pass


@dataclass
class JobOutcome:
    job_id: str
    rounds: int
    workers: List[str]
    parameter_server: str
    metrics: List[Dict[str, Any]]
    finished: bool
    removed_workers: List[str] = field(default_factory=list)   # train workers dropped after a lost lease
    joined_workers: List[str] = field(default_factory=list)    # replacements admitted into the running job


def free_tcp_port() -> int:
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as sk:
        sk.bind(("127.0.0.1", 0))
        return int(sk.getsockname()[1])


def rendezvous_series(first: str):
    """Addresses for the successive process groups of one collective job (a new generation forms after every
    membership change).  On one box each is a free loopback port.  With a configured ``rendezvous`` on another host
    (a team across boxes: the store is hosted by rank 0 THERE, so this process cannot probe for free ports) the host is
    kept and the port counts up from the configured one -- reserve a small range."""
    from urllib.parse import urlparse

    u = urlparse(first)
    if u.hostname in (None, "127.0.0.1", "localhost", "::1"):
        return lambda: f"tcp://127.0.0.1:{free_tcp_port()}"
    state = {"port": int(u.port or 29400)}

    def nxt() -> str:
        state["port"] += 1
        host = f"[{u.hostname}]" if ":" in u.hostname else u.hostname
        return f"tcp://{host}:{state['port']}"
    return nxt


def worker_batch_size(offered_gpu: float, required_gpu: float, max_batch: Optional[int]) -> int:
    """``floor(offered.gpu / required.gpu).min(max_batch_size)`` (hypha-scheduler.rs:320-322)."""
    bs = int(offered_gpu // required_gpu) if required_gpu > 0 else 1
    bs = max(1, bs)
    return min(bs, max_batch) if max_batch else bs


class _Tee(MetricsConnector):
    """Keeps a copy of every metric for the JobOutcome while forwarding to the configured sink."""

    def __init__(self, inner: MetricsConnector, collected: List[Dict[str, Any]]):
        self.inner, self.collected = inner, collected

    async def send(self, worker_id, round, name, value):
        self.collected.append({"worker_id": worker_id, "round": round, "metric_name": name, "value": value})
        await self.inner.send(worker_id, round, name, value)


class DiLoCoJobRun:
    """Owns the handles, services and tasks of one job; ``run()`` returns its :class:`JobOutcome`."""

    def __init__(self, net: Network, job, *, telemetry, stop: asyncio.Event, data_provider,
                 metrics_connector: Optional[MetricsConnector] = None, status_bridge: Optional[str] = None,
                 allocation_pause_s: float = ALLOCATION_PAUSE_S):
        self.net, self.job, self.res = net, job, job.resources
        self.telemetry, self.log, self.stop = telemetry, telemetry.logger, stop
        self.get_data_provider = data_provider
        self.sink = metrics_connector or (AimConnector(status_bridge) if status_bridge else NoOpConnector())
        if telemetry is not None and getattr(telemetry, "enabled", False):
            # an OTLP endpoint is configured: training metrics go there as well (hypha.job.* gauges)
            from .scheduling import FanOutConnector, OtelConnector

            self.sink = FanOutConnector(self.sink, OtelConnector(telemetry))
        self.allocation_pause_s = allocation_pause_s
        self.collective = job.aggregation.mode == "collective"
        self.allocator = GreedyWorkerAllocator(net)
        self.job_id = M.new_id()
        self.handles: List[WorkerHandle] = []
        self.services: List[Any] = []
        self.tasks: List[Task] = []
        self.background: List[asyncio.Future] = []
        self.workers: List[WorkerHandle] = []
        self.ps: Optional[WorkerHandle] = None
        self.batch: Optional[BatchSchedulerService] = None
        self.data_svc: Optional[DataSchedulerService] = None
        self.live: List[str] = []
        self.removed: List[str] = []
        self.joined: List[str] = []
        self.collected: List[Dict[str, Any]] = []
        self._statuses: asyncio.Queue = asyncio.Queue(64)      # (peer, JobStatus value) from every task
        self._replacements = 0
        to_res = lambda r: M.Resources(gpu=r.gpu, cpu=r.cpu, storage=r.storage, memory=r.memory)
        self.wspec = M.WorkerSpec(to_res(self.res.worker), [M.ExecutorDescriptor("train", M.TRAIN_EXECUTOR_NAME)])
        self.pspec = M.WorkerSpec(to_res(self.res.parameter_server),
                                  [M.ExecutorDescriptor("aggregate", M.AGGREGATE_EXECUTOR_NAME)])

    # -- allocate -------------------------------------------------------------------------------
    async def allocate(self) -> None:
        res = self.res
        with self.telemetry.span("allocate.workers", num=res.num_workers):
            self.workers = await self.allocator.request(self.wspec, res.worker_price.bid, res.worker_price.max,
                                                        res.num_workers)
        self.handles += self.workers
        if self.collective:
            # B200 path: no aggregate worker.  The train workers run the outer step among themselves
            # (fused peer-memory kernel on GPUs / gloo on CPU) and report `updated` for the group.
            if len(self.workers) < res.num_workers:
                raise JobError(f"allocation incomplete: {len(self.workers)}/{res.num_workers} workers")
            self.ps = self.workers[0]
            return
        await asyncio.sleep(self.allocation_pause_s)
        with self.telemetry.span("allocate.parameter_server"):
            ps_list = await self.allocator.request(self.pspec, res.parameter_server_price.bid,
                                                   res.parameter_server_price.max, 1)
        self.handles += ps_list
        if len(self.workers) < res.num_workers or not ps_list:
            raise JobError(f"allocation incomplete: {len(self.workers)}/{res.num_workers} workers, {len(ps_list)} PS")
        self.ps = ps_list[0]

    # -- dispatch -------------------------------------------------------------------------------
    def _job_messages(self) -> None:
        job = self.job
        self.model = M.Model(job.model.type, M.Fetch.huggingface(job.model.repository, job.model.revision,
                                                                 job.model.filenames, job.model.token),
                             list(job.model.input_names))
        self.pre = None
        if job.preprocessor is not None:
            p = job.preprocessor
            self.pre = M.Preprocessor(p.type, M.Fetch.huggingface(p.repository, p.revision, p.filenames, p.token),
                                      list(p.input_names))
        self.adam = M.Adam(job.inner_optimizer.learning_rate,
                           tuple(job.inner_optimizer.betas) if job.inner_optimizer.betas else None,
                           job.inner_optimizer.epsilon)
        self.nesterov = M.Nesterov(job.outer_optimizer.learning_rate, job.outer_optimizer.momentum)

    def train_config(self, w: WorkerHandle) -> M.TrainExecutorConfig:
        bs = worker_batch_size(w.resources.gpu, self.res.worker.gpu, self.job.rounds.max_batch_size)
        self.batch.add_worker(w.peer_id, bs)
        return M.TrainExecutorConfig(model=self.model, data=M.Fetch.scheduler(self.net.peer_id, self.job.dataset.dataset),
                                     updates=M.Send.to_peers([self.ps.peer_id]),
                                     results=M.Receive.from_peers([self.ps.peer_id]),
                                     optimizer=self.adam, batch_size=bs, preprocessor=self.pre,
                                     weighting=self.job.aggregation.weighting,
                                     precision=self.job.precision.gemm,
                                     accumulation={"micro_batches": self.job.precision.micro_batches,
                                                   "main_grad": self.job.precision.main_grad})

    async def _dispatch_train(self, w: WorkerHandle, tcfg: M.TrainExecutorConfig) -> None:
        spec = M.JobSpec(self.job_id, M.Executor.train(M.TRAIN_EXECUTOR_NAME, tcfg))
        self._watch_task(await Task.try_new(self.net, spec, [w]))

    def _watch_task(self, task: Task) -> None:
        """Forward the task's JobStatus reports (task.rs:27-121) into the watch loop."""
        self.tasks.append(task)

        async def pump():
            while True:
                peer, status = await task.statuses.get()
                await self._statuses.put((peer, status))

        self.background.append(asyncio.ensure_future(pump()))

    async def dispatch(self) -> None:
        job = self.job
        provider, record = await self.get_data_provider(self.net, job.dataset.dataset)
        self.data_svc = DataSchedulerService(self.net, job.dataset.dataset, provider, record.num_slices)
        self.data_svc.start()
        self.services.append(self.data_svc)
        self.batch = BatchSchedulerService(self.net, self.job_id, self.ps.peer_id,
                                           job.rounds.avg_samples_between_updates, job.rounds.update_rounds)
        self.batch.start()
        self.services.append(self.batch)
        self._job_messages()
        self.live = [w.peer_id for w in self.workers]
        ck = job.checkpoint
        resume_round = 0
        if ck.dir and ck.resume:
            from ..worker.jobs import latest_checkpoint

            found = latest_checkpoint(ck.dir)
            want = "collective" if self.collective else "parameter-server"
            if found is not None and found.get("kind") != want:
                raise JobError(f"the checkpoint in {ck.dir} was written in {found.get('kind')} mode, "
                               f"this job aggregates in {want} mode")
            if found is not None:
                resume_round = min(found["round"], job.rounds.update_rounds)
                self.batch.start_at_round(resume_round)
                self.log.info("resuming job from the checkpoint of round %d in %s", resume_round, ck.dir)
            else:
                self.log.warning("checkpoint.resume is set but %s holds no checkpoint: starting at round 0", ck.dir)
        rendezvous = job.aggregation.rendezvous or f"tcp://127.0.0.1:{free_tcp_port()}"
        if self.collective:
            self.batch.enable_collective(self.live, rendezvous, rendezvous_series(rendezvous))
        for i, w in enumerate(self.workers):
            tcfg = self.train_config(w)
            if self.collective:
                tcfg.collective = {"rank": i, "world": len(self.workers), "rendezvous": rendezvous,
                                   "outer": self.nesterov.to_wire(), "peer": w.peer_id, "peers": list(self.live)}
                if ck.dir:      # rank 0 of the current group writes theta_prev + outer momentum every `every` rounds
                    tcfg.collective["checkpoint"] = {"dir": ck.dir, "every": ck.every, "resume_round": resume_round}
            elif resume_round:
                tcfg.join = {"after_round": resume_round, "resume": True}      # theta = theta_0(seed) + U
            await self._dispatch_train(w, tcfg)
        if not self.collective:
            acfg = M.AggregateExecutorConfig(updates=M.Receive.from_peers(self.live), results=M.Send.to_peers(self.live),
                                             optimizer=self.nesterov)
            if ck.dir:
                acfg.checkpoint = {"dir": ck.dir, "every": ck.every}
                if resume_round:
                    acfg.checkpoint["resume_round"] = resume_round
            spec = M.JobSpec(self.job_id, M.Executor.aggregate(M.AGGREGATE_EXECUTOR_NAME, acfg))
            self._watch_task(await Task.try_new(self.net, spec, [self.ps]))
        if ck.dir:
            if resume_round:
                self._restore_scheduler_state(ck.dir, resume_round)
            self.background.append(asyncio.ensure_future(self._checkpoint_scheduler_state(ck.dir, ck.every)))

    # -- scheduler side of the job checkpoint -----------------------------------------------------
    def _restore_scheduler_state(self, ckpt_dir: str, round_no: int) -> None:
        """The aggregate worker's checkpoint holds the outer state; the scheduler adds its own: which slices
        of the current data epoch were already handed out (round + slice-tracker state)."""
        path = os.path.join(ckpt_dir, "scheduler_state.json")
        try:
            st = json.load(open(path))
            if int(st["round"]) != int(round_no):
                self.log.info("scheduler state in %s is of round %s, resuming round %d: new data epoch", path, st["round"], round_no)
                return
            self.data_svc.tracker.restore([(str(o), bool(p)) for o, p in st["slices"]], int(st["epochs"]))
            self.log.info("data scheduler resumed: %d of %d slices of epoch %d already used",
                          sum(1 for _, p in st["slices"] if p), len(st["slices"]), int(st["epochs"]))
        except FileNotFoundError:
            pass
        except (OSError, ValueError, KeyError, TypeError) as exc:
            self.log.warning("ignoring unreadable scheduler state %s: %s", path, exc)

    def _write_scheduler_state(self, ckpt_dir: str, rnd: int) -> None:
        try:
            os.makedirs(ckpt_dir, exist_ok=True)
            tmp = os.path.join(ckpt_dir, "scheduler_state.json.tmp")
            with open(tmp, "w") as f:
                json.dump({"round": rnd, "epochs": int(self.data_svc.tracker.rounds()),
                           "slices": [[o, bool(p)] for o, p in self.data_svc.tracker.snapshot()]}, f)
            os.replace(tmp, os.path.join(ckpt_dir, "scheduler_state.json"))
        except OSError as exc:
            self.log.warning("could not write the scheduler state: %s", exc)

    async def _checkpoint_scheduler_state(self, ckpt_dir: str, every: int) -> None:
        last = self.batch.round
        while True:
            finished = self.batch.done.is_set()
            rnd = self.batch.round
            if rnd != last and (finished or rnd % max(1, int(every)) == 0):
                last = rnd
                self._write_scheduler_state(ckpt_dir, rnd)
            if finished:
                return
            await asyncio.sleep(0.25)

    # -- elasticity -----------------------------------------------------------------------------
    async def _tell_aggregate_worker(self) -> None:
        await self.net.request(self.ps.peer_id, M.API_PROTOCOL, M.ApiRequest.job_membership(self.job_id, self.live),
                               timeout=5.0)

    async def _drop_worker(self, peer: str, why: str = "lost the lease of") -> None:
        self.live.remove(peer)
        self.removed.append(peer)
        self.log.warning("%s train worker %s: continuing with %d worker(s)", why, peer, len(self.live))
        self.batch.remove_worker(peer)
        self.data_svc.remove_worker(peer)
        if not self.collective:
            try:
                await self._tell_aggregate_worker()
            except NetworkError as exc:
                self.log.warning("could not tell the aggregate worker about the new membership: %s", exc)
        if self.job.elastic.replace and self._replacements < self.job.elastic.max_replacements:
            self._replacements += 1
            self.background.append(asyncio.ensure_future(self._replace_worker(peer)))
        elif self.job.elastic.replace:
            self.log.warning("not replacing %s: %d replacements already made", peer, self._replacements)

    async def _replace_worker(self, lost: str) -> None:
        """Lease one more train worker and let it join the running job (parameter-server mode): it is
        dispatched with `join`, and the aggregate worker -- told the new membership -- sends it the
        cumulative update at the next round boundary."""
        try:
            if self.batch.round + 2 > self.job.rounds.update_rounds:
                self.log.info("not replacing %s: fewer than two rounds left", lost)
                return
            new: List[WorkerHandle] = []
            for attempt in range(REPLACE_ATTEMPTS):
                # a worker whose executor crashed still holds its old lease until it expires (10 s), so
                # the first advertisement may go unanswered
                try:
                    got = await self.allocator.request(self.wspec, self.res.worker_price.bid,
                                                       self.res.worker_price.max, 1)
                except Exception as exc:                      # NoOffersReceived and friends
                    self.log.info("replacement attempt %d for %s: %s", attempt + 1, lost, exc)
                    got = []
                new = [h for h in got if h.peer_id not in self.live and (self.collective or h.peer_id != self.ps.peer_id)]
                for h in got:
                    if h not in new:
                        await h.release()
                if new or self.batch.done.is_set() or self.batch.round + 2 > self.job.rounds.update_rounds:
                    break
                await asyncio.sleep(REPLACE_RETRY_S)
            if not new:
                self.log.warning("no replacement offered for %s", lost)
                return
            h = new[0]
            self.handles.append(h)
            tcfg = self.train_config(h)
            if self.collective:
                # joins at the next `update` barrier: the survivors re-form their group with it and rank 0
                # broadcasts the current model and outer momentum (CPU engine; see executor/training.py)
                # `slot`: the rank number it takes over (engines that shard state by rank: the fused GPU step)
                tcfg.collective = {"rank": -1, "world": len(self.live) + 1, "rendezvous": None,
                                   "outer": self.nesterov.to_wire(), "peer": h.peer_id,
                                   "peers": [w.peer_id for w in self.workers], "join": True,
                                   "after_round": self.batch.round, "slot": self.batch.reserve_slot(h.peer_id),
                                   "slots_world": len(self.workers)}
            else:
                tcfg.join = {"after_round": self.batch.round, "replaces": lost}
            try:
                await self._dispatch_train(h, tcfg)
            except BaseException:
                self.batch.release_slot(h.peer_id)
                raise
            self.live.append(h.peer_id)
            self.joined.append(h.peer_id)
            if self.collective:
                self.batch.add_member(h.peer_id)
            else:
                await self._tell_aggregate_worker()
            if h.task is not None:
                self._waiters[h.task] = h
                self._wake.set()
            self.log.info("replacement %s joins job %s after round %d", h.peer_id, self.job_id[:8], self.batch.round)
        except (NetworkError, TaskError, JobError, asyncio.TimeoutError) as exc:
            self.log.warning("could not replace %s: %s", lost, exc)

    # -- watch ----------------------------------------------------------------------------------
    async def watch(self) -> str:
        """Wait for: job done | stop requested | lease lost.  The reference exits on the first lost lease
        (hypha-scheduler.rs:396-412); with `elastic.enabled` a lost TRAIN worker is removed from the
        batch/slice schedulers, the aggregate job is told the new membership, and the job goes on while
        at least `elastic.min_workers` remain.  Returns the reason the wait ended."""
        el = self.job.elastic
        self._waiters: Dict[Any, Any] = {asyncio.ensure_future(self.batch.done.wait()): "done",
                                         asyncio.ensure_future(self.stop.wait()): "stop"}
        for h in self.handles:
            if h.task is not None:
                self._waiters[h.task] = h
        self._wake = asyncio.Event()        # set when a replacement's lease task must be watched too

        def can_drop(peer: str) -> bool:
            return el.enabled and peer in self.live and len(self.live) - 1 >= max(1, el.min_workers)

        while True:
            wake_task = asyncio.ensure_future(self._wake.wait())
            status_task = asyncio.ensure_future(self._statuses.get())
            done, _ = await asyncio.wait(set(self._waiters.keys()) | {wake_task, status_task},
                                         return_when=asyncio.FIRST_COMPLETED)
            if wake_task in done:
                self._wake.clear()
            else:
                wake_task.cancel()
            if status_task in done:
                peer, status = status_task.result()
                if status == M.JobStatus.FAILED.value and not self.batch.done.is_set():
                    # the executor died although the worker (and its lease) is alive
                    if not self.collective and peer == self.ps.peer_id:
                        return f"aggregate-job-{status.lower()}"
                    if peer in self.live:
                        if not can_drop(peer):
                            return f"job-{status.lower()}:{peer}"
                        handle = next((h for h in self.handles if h.peer_id == peer), None)
                        if handle is not None:
                            if handle.task is not None:
                                self._waiters.pop(handle.task, None)
                            await handle.release()          # free the worker: it may offer again
                            self.handles.remove(handle)
                        await self._drop_worker(peer, why=f"job {status.lower()} on")
            else:
                status_task.cancel()
            done = done - {wake_task, status_task}
            if not done:
                continue
            what = self._waiters.pop(next(iter(done)), None)
            if what is None:
                continue
            if isinstance(what, str):
                return what
            peer = what.peer_id
            if not can_drop(peer):
                return f"lease:{peer}"
            await self._drop_worker(peer)

    # -- run ------------------------------------------------------------------------------------
    async def run(self) -> JobOutcome:
        try:
            await self.allocate()
            await self.dispatch()
            bridge = MetricsBridge(_Tee(self.sink, self.collected))
            bridge.register_stream(self.batch.metrics)
            bridge_task = asyncio.ensure_future(bridge.run())
            self.background.append(bridge_task)
            reason = await self.watch()
            finished = self.batch.done.is_set()
            if not finished:
                self.log.warning("scheduler stopping before completion: %s", reason)
            await asyncio.sleep(0.05)              # let the last metrics drain
            return JobOutcome(self.job_id, self.batch.round, [w.peer_id for w in self.workers],
                              "collective" if self.collective else self.ps.peer_id, self.collected, finished,
                              self.removed, self.joined)
        finally:
            await self.close()

    async def close(self) -> None:
        ck = self.job.checkpoint
        if ck.dir and getattr(self, "data_svc", None) is not None and self.batch.round > 0:
            self._write_scheduler_state(ck.dir, self.batch.round)      # the state the last checkpoint belongs to
        for b in self.background:
            if not b.done():
                b.cancel()
        for t in self.tasks:
            t.close()
        for s in self.services:
            s.stop()
        for h in self.handles:
            await h.release()
