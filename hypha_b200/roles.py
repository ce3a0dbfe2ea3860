"""The four long-running roles -- gateway, data node, worker, scheduler -- as asyncio programs.

Reference main flows: crates/gateway/src/bin/hypha-gateway.rs:36-164,
crates/data/src/bin/hypha-data.rs:33-233, crates/worker/src/bin/hypha-worker.rs:42-283,
crates/scheduler/src/bin/hypha-scheduler.rs:54-432.  Common start-up sequence for every role:
telemetry -> Network -> listen -> external addresses -> dial gateways -> DHT bootstrap ->
health handler -> role logic -> graceful shutdown.

Roles run either all in one process over ``memory://`` links (the north-star layout: the 8 GPUs
of a box are the workers) or as separate processes over ``unix://`` / ``tcp://``.
"""
from __future__ import annotations

import asyncio
import json
import math
import os
import time
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional

from . import config as C
from . import messages as M
from .scheduler import (AimConnector, BatchSchedulerService, Connector as MetricsConnector, DataSchedulerService,
                        GreedyWorkerAllocator, JsonlConnector, MetricsBridge, NoOpConnector, Task, WorkerHandle)
from .scheduler.allocator import TaskError
from .telemetry import Telemetry
from .transport import Network, NetworkError, RecordNotFound
from .worker import Arbiter, Connector, JobManager

ALLOCATION_PAUSE_S = 1.0          # hypha-scheduler.rs:241 (let losers' tentative leases expire)


class RoleError(Exception):
    pass


@dataclass
class Node:
    """Started network + telemetry of one role."""

    name: str
    network: Network
    telemetry: Telemetry
    ready: asyncio.Event = field(default_factory=asyncio.Event)
    stop: asyncio.Event = field(default_factory=asyncio.Event)

    @property
    def peer_id(self) -> str:
        return self.network.peer_id

    async def shutdown(self) -> None:
        self.stop.set()
        await self.network.close()
        self.telemetry.shutdown()


async def start_node(role: str, cfg: Any, *, peer_id: Optional[str] = None, relay_server: bool = False,
                     dht_server: bool = False, need_bootstrap: bool = True) -> Node:
    tel = Telemetry.from_config(f"hypha-{role}", cfg)
    C.load_tls(cfg)                         # validates that the configured PEM files are readable
    tls = {k: getattr(cfg, k, "") for k in ("cert_pem", "key_pem", "trust_pem", "crls_pem")}
    if peer_id is None and getattr(cfg, "key_pem", ""):
        from .transport.utils import peer_id_from_key_pem

        peer_id = peer_id_from_key_pem(cfg.key_pem)      # identity = public key, like the reference
    net = Network(peer_id or f"{role}-{M.new_id()[:8]}", agent=f"hypha-{role}", relay_server=relay_server,
                  dht_server=dht_server, tls=tls)          # mTLS applies to tcp:// links only
    if any(a.startswith("tcp://") for a in cfg.listen_addresses) and net.tls is None:
        tel.logger.warning("tcp:// listener without cert_pem/key_pem/trust_pem: the link is NOT encrypted")
    tel.register_bandwidth(net.counters)
    node = Node(role, net, tel)
    net.enable_dht_handlers()
    for addr in cfg.listen_addresses:
        await net.listen(addr)
    for addr in getattr(cfg, "external_addresses", []) or []:
        net.add_external_address(addr)

    async def health(peer, msg):            # /hypha-health/0.0.1: healthy == ready
        return M.HealthResponse(node.ready.is_set())

    net.on(M.HEALTH_PROTOCOL, None, health)
    gateways = getattr(cfg, "gateway_addresses", []) or []
    dialed = 0
    for addr in gateways:
        try:
            await net.dial(addr)
            dialed += 1
        except NetworkError as exc:
            tel.logger.warning("dial gateway %s failed: %s", addr, exc)
    if gateways and dialed == 0:
        raise RoleError(f"{role}: none of the gateways {gateways} could be dialed")
    if need_bootstrap and gateways:
        await net.wait_for_bootstrap(10.0)
    return node


async def probe(addr: str, timeout: float = 5.0, tls: Optional[dict] = None) -> bool:
    """``probe <addr>``: dial, send ``health::Request``, true iff healthy
    (hypha-scheduler.rs:494-535)."""
    net = Network(f"probe-{M.new_id()[:8]}", tls=tls)
    try:
        peer = await net.dial(addr, timeout=timeout)
        resp = await net.request(peer, M.HEALTH_PROTOCOL, M.HealthRequest(), timeout=timeout)
        return bool(M.HealthResponse.from_wire(resp).healthy)
    except NetworkError:
        return False
    finally:
        await net.close()


# ---------------------------------------------------------------------------------------------
# gateway
# ---------------------------------------------------------------------------------------------
async def run_gateway(cfg: C.GatewayConfig, *, peer_id: Optional[str] = None,
                      started: Optional[asyncio.Future] = None) -> None:
    node = await start_node("gateway", cfg, peer_id=peer_id, relay_server=True, dht_server=True,
                            need_bootstrap=False)
    node.ready.set()                        # gateway readiness == listening (hypha-gateway.rs:78,121)
    node.telemetry.logger.info("gateway %s listening on %s", node.peer_id, node.network.listen_addrs)
    if started is not None and not started.done():
        started.set_result(node)
    try:
        await node.stop.wait()
    finally:
        await node.shutdown()


# ---------------------------------------------------------------------------------------------
# data node
# ---------------------------------------------------------------------------------------------
def dataset_files(path: str) -> List[str]:
    """Slices of a dataset directory in SORTED order (the reference indexes ``read_dir()`` order,
    which is unspecified: quirk Q5)."""
    return sorted(os.path.join(path, f) for f in os.listdir(path) if os.path.isfile(os.path.join(path, f)))


async def run_data(cfg: C.DataConfig, *, peer_id: Optional[str] = None,
                   started: Optional[asyncio.Future] = None) -> None:
    C.validate(cfg)
    node = await start_node("data", cfg, peer_id=peer_id)
    name = os.path.basename(os.path.normpath(cfg.dataset_path))
    files = dataset_files(cfg.dataset_path)
    if not files:
        raise RoleError(f"dataset directory {cfg.dataset_path} is empty")

    async def serve(peer: str, header: Any):
        sl = M.DataSlice.from_wire(header)
        if sl.dataset != name:
            raise RoleError(f"unknown dataset {sl.dataset!r}")
        if not (0 <= sl.index < len(files)):            # correct bounds (reference uses `>`: Q5)
            raise RoleError(f"slice {sl.index} out of range 0..{len(files) - 1}")
        return files[sl.index]

    node.network.serve_pull(serve)
    await node.network.put_record(name, json.dumps(M.DataRecord(len(files)).to_wire()).encode())
    await node.network.provide(name)
    node.ready.set()
    node.telemetry.logger.info("data node %s announces %s (%d slices)", node.peer_id, name, len(files))
    if started is not None and not started.done():
        started.set_result(node)
    try:
        await node.stop.wait()
    finally:
        await node.shutdown()


# ---------------------------------------------------------------------------------------------
# worker
# ---------------------------------------------------------------------------------------------
async def run_worker(cfg: C.WorkerConfig, *, peer_id: Optional[str] = None,
                     started: Optional[asyncio.Future] = None, log: Callable[[str], None] = print) -> None:
    C.validate(cfg)
    node = await start_node("worker", cfg, peer_id=peer_id)
    net = node.network
    connector = Connector(net)
    jobs = JobManager(connector, net, cfg.work_dir, cfg.executors, log=log)
    cap = M.Resources(gpu=cfg.resources.gpu, cpu=cfg.resources.cpu, storage=cfg.resources.storage,
                      memory=cfg.resources.memory)
    arbiter = Arbiter(net, cap, jobs, price=cfg.offer.price, floor=cfg.offer.floor,
                      strategy=cfg.offer.strategy.type, log=log)
    node.arbiter = arbiter  # type: ignore[attr-defined]
    node.jobs = jobs  # type: ignore[attr-defined]
    node.ready.set()                         # listening + bootstrapped (hypha-worker.rs:84-87,200)
    if started is not None and not started.done():
        started.set_result(node)
    try:
        await arbiter.run(node.stop)
    finally:
        await jobs.shutdown()
        await node.shutdown()


# ---------------------------------------------------------------------------------------------
# scheduler
# ---------------------------------------------------------------------------------------------
class _TaskHandle:
    """Lets a background asyncio task sit in the scheduler's `services` list (stopped on exit)."""

    def __init__(self, task: "asyncio.Future"):
        self.task = task

    def stop(self) -> None:
        if not self.task.done():
            self.task.cancel()


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


def _free_tcp_port() -> int:
    import socket

    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as sk:
        sk.bind(("127.0.0.1", 0))
        return int(sk.getsockname()[1])


def worker_batch_size(offered_gpu: float, required_gpu: float, max_batch: Optional[int]) -> int:
    """``floor(offered.gpu / required.gpu).min(max_batch_size)`` (hypha-scheduler.rs:320-322)."""
    bs = int(math.floor(offered_gpu / required_gpu)) if required_gpu > 0 else 1
    bs = max(bs, 1)
    return min(bs, max_batch) if max_batch else bs


async def get_data_provider(net: Network, dataset: str, retries: int = 20) -> tuple:
    for _ in range(retries):
        try:
            value, publisher = await net.get_record(dataset)
            return publisher, M.DataRecord.from_wire(json.loads(value))
        except RecordNotFound:
            await asyncio.sleep(0.25)
    raise RoleError(f"dataset {dataset!r} is not announced by any data node")


async def run_scheduler(cfg: C.SchedulerConfig, *, peer_id: Optional[str] = None,
                        started: Optional[asyncio.Future] = None,
                        metrics_connector: Optional[MetricsConnector] = None,
                        allocation_pause_s: float = ALLOCATION_PAUSE_S) -> JobOutcome:
    C.validate(cfg)
    node = await start_node("scheduler", cfg, peer_id=peer_id)
    net, log = node.network, node.telemetry.logger
    job = cfg.scheduler.job
    res = job.resources
    node.ready.set()
    if started is not None and not started.done():
        started.set_result(node)
    handles: List[WorkerHandle] = []
    services: List[Any] = []
    tasks: List[Task] = []
    try:
        allocator = GreedyWorkerAllocator(net)
        to_res = lambda r: M.Resources(gpu=r.gpu, cpu=r.cpu, storage=r.storage, memory=r.memory)
        wspec = M.WorkerSpec(to_res(res.worker), [M.ExecutorDescriptor("train", M.TRAIN_EXECUTOR_NAME)])
        with node.telemetry.span("allocate.workers", num=res.num_workers):
            workers = await allocator.request(wspec, res.worker_price.bid, res.worker_price.max, res.num_workers)
        handles += workers
        collective = job.aggregation.mode == "collective"
        if collective:
            # B200 path: no aggregate worker.  The train workers run the outer step among themselves
            # (fused peer-memory kernel on GPUs / gloo on CPU); rank 0 reports `updated` for the group.
            if len(workers) < res.num_workers:
                raise RoleError(f"allocation incomplete: {len(workers)}/{res.num_workers} workers")
            ps = workers[0]
        else:
            await asyncio.sleep(allocation_pause_s)
            pspec = M.WorkerSpec(to_res(res.parameter_server), [M.ExecutorDescriptor("aggregate", M.AGGREGATE_EXECUTOR_NAME)])
            with node.telemetry.span("allocate.parameter_server"):
                ps_list = await allocator.request(pspec, res.parameter_server_price.bid, res.parameter_server_price.max, 1)
            handles += ps_list
            if len(workers) < res.num_workers or not ps_list:
                raise RoleError(f"allocation incomplete: {len(workers)}/{res.num_workers} workers, {len(ps_list)} PS")
            ps = ps_list[0]
        provider, record = await get_data_provider(net, job.dataset.dataset)
        data_svc = DataSchedulerService(net, job.dataset.dataset, provider, record.num_slices)
        data_svc.start()
        services.append(data_svc)

        job_id = M.new_id()
        batch = BatchSchedulerService(net, job_id, ps.peer_id, job.rounds.avg_samples_between_updates,
                                      job.rounds.update_rounds)
        batch.start()
        services.append(batch)
        worker_ids = [w.peer_id for w in workers]
        model = M.Model(job.model.type, M.Fetch.huggingface(job.model.repository, job.model.revision,
                                                            job.model.filenames, job.model.token),
                        list(job.model.input_names))
        pre = None
        if job.preprocessor is not None:
            p = job.preprocessor
            pre = M.Preprocessor(p.type, M.Fetch.huggingface(p.repository, p.revision, p.filenames, p.token),
                                 list(p.input_names))
        adam = M.Adam(job.inner_optimizer.learning_rate,
                      tuple(job.inner_optimizer.betas) if job.inner_optimizer.betas else None,
                      job.inner_optimizer.epsilon)
        nesterov = M.Nesterov(job.outer_optimizer.learning_rate, job.outer_optimizer.momentum)
        rendezvous = job.aggregation.rendezvous or f"tcp://127.0.0.1:{_free_tcp_port()}"

        def train_config(w: WorkerHandle) -> M.TrainExecutorConfig:
            bs = worker_batch_size(w.resources.gpu, res.worker.gpu, job.rounds.max_batch_size)
            batch.add_worker(w.peer_id, bs)
            return M.TrainExecutorConfig(model=model, data=M.Fetch.scheduler(net.peer_id, job.dataset.dataset),
                                         updates=M.Send.to_peers([ps.peer_id]), results=M.Receive.from_peers([ps.peer_id]),
                                         optimizer=adam, batch_size=bs, preprocessor=pre)

        for i, w in enumerate(workers):
            tcfg = train_config(w)
            if collective:
                tcfg.collective = {"rank": i, "world": len(workers), "rendezvous": rendezvous,
                                   "outer": nesterov.to_wire()}
            spec = M.JobSpec(job_id, M.Executor.train(M.TRAIN_EXECUTOR_NAME, tcfg))
            tasks.append(await Task.try_new(net, spec, [w]))
        if not collective:
            acfg = M.AggregateExecutorConfig(updates=M.Receive.from_peers(worker_ids), results=M.Send.to_peers(worker_ids),
                                             optimizer=nesterov)
            tasks.append(await Task.try_new(net, M.JobSpec(job_id, M.Executor.aggregate(M.AGGREGATE_EXECUTOR_NAME, acfg)), [ps]))

        collected: List[Dict[str, Any]] = []

        class _Tee(MetricsConnector):
            def __init__(self, inner):
                self.inner = inner

            async def send(self, worker_id, round, name, value):
                collected.append({"worker_id": worker_id, "round": round, "metric_name": name, "value": value})
                await self.inner.send(worker_id, round, name, value)

        sink = metrics_connector or (AimConnector(cfg.status_bridge) if cfg.status_bridge else NoOpConnector())
        bridge = MetricsBridge(_Tee(sink))
        bridge.register_stream(batch.metrics)
        bridge_task = asyncio.ensure_future(bridge.run())

        # wait: job done | stop requested | lease lost.  The reference exits on the first lost lease
        # (hypha-scheduler.rs:396-412, quirk Q9); with `elastic.enabled` a lost TRAIN worker is removed
        # from the batch/slice schedulers, the aggregate job is told the new membership, and the job
        # goes on while at least `elastic.min_workers` remain.
        waiters: Dict[Any, Any] = {asyncio.ensure_future(batch.done.wait()): "done",
                                   asyncio.ensure_future(node.stop.wait()): "stop"}
        for h in handles:
            if h.task is not None:
                waiters[h.task] = h
        live = list(worker_ids)
        removed: List[str] = []
        joined: List[str] = []
        reason = "done"
        wake = asyncio.Event()                      # set when a replacement's lease task must be watched too

        async def replace_worker(lost: str) -> None:
            """Lease one more train worker and let it join the running job (parameter-server mode): it
            is dispatched with `join`, and the aggregate worker -- told the new membership -- sends it the
            cumulative update at the next round boundary."""
            try:
                if batch.round + 2 > job.rounds.update_rounds:
                    log.info("not replacing %s: fewer than two rounds left", lost)
                    return
                new = await allocator.request(wspec, res.worker_price.bid, res.worker_price.max, 1)
                new = [h for h in new if h.peer_id not in live and h.peer_id != ps.peer_id]
                if not new:
                    log.warning("no replacement offered for %s", lost)
                    return
                h = new[0]
                handles.append(h)
                tcfg = train_config(h)
                tcfg.join = {"after_round": batch.round, "replaces": lost}
                tasks.append(await Task.try_new(net, M.JobSpec(job_id, M.Executor.train(M.TRAIN_EXECUTOR_NAME, tcfg)), [h]))
                live.append(h.peer_id)
                joined.append(h.peer_id)
                await net.request(ps.peer_id, M.API_PROTOCOL, M.ApiRequest.job_membership(job_id, live), timeout=5.0)
                if h.task is not None:
                    waiters[h.task] = h
                    wake.set()
                log.info("replacement %s joins job %s after round %d", h.peer_id, job_id[:8], batch.round)
            except (NetworkError, TaskError, RoleError, asyncio.TimeoutError) as exc:
                log.warning("could not replace %s: %s", lost, exc)

        while True:
            wake_task = asyncio.ensure_future(wake.wait())
            done, _ = await asyncio.wait(set(waiters.keys()) | {wake_task}, return_when=asyncio.FIRST_COMPLETED)
            if wake_task in done:
                wake.clear()
                if len(done) == 1:
                    continue
                done = done - {wake_task}
            else:
                wake_task.cancel()
            first = next(iter(done))
            what = waiters.pop(first)
            if isinstance(what, str):
                reason = what
                break
            peer = what.peer_id
            reason = f"lease:{peer}"
            if not (job.elastic.enabled and peer in live and len(live) - 1 >= max(1, job.elastic.min_workers)):
                break
            live.remove(peer)
            removed.append(peer)
            log.warning("lost the lease of train worker %s: continuing with %d worker(s)", peer, len(live))
            batch.remove_worker(peer)
            data_svc.remove_worker(peer)
            if not collective:
                try:
                    await net.request(ps.peer_id, M.API_PROTOCOL, M.ApiRequest.job_membership(job_id, live), timeout=5.0)
                except NetworkError as exc:
                    log.warning("could not tell the aggregate worker about the new membership: %s", exc)
                if job.elastic.replace:
                    services.append(_TaskHandle(asyncio.ensure_future(replace_worker(peer))))
        finished = batch.done.is_set()
        if not finished:
            log.warning("scheduler stopping before completion: %s", reason)
        await asyncio.sleep(0.05)              # let the last metrics drain
        bridge_task.cancel()
        return JobOutcome(job_id, batch.round, worker_ids, "collective" if collective else ps.peer_id, collected,
                          finished, removed, joined)
    finally:
        for t in tasks:
            t.close()
        for s in services:
            s.stop()
        for h in handles:
            await h.release()
        await node.shutdown()
