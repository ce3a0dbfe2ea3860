"""Batch scheduler (progress state machine), data scheduler (slice hand-out), metrics bridge.

Reference: crates/scheduler/src/scheduling/batch_scheduler.rs:54-220,
crates/scheduler/src/scheduling/data_scheduler.rs:28-103, crates/scheduler/src/metrics_bridge.rs:18-146.
The state machine, the sync-point simulation and the slice tracker are the C++ core; these
classes bind them to the request/response fabric.
"""
from __future__ import annotations

import asyncio
import json
import time
from typing import Any, Callable, Dict, List, Optional

from .. import messages as M
from ..ops.native import load_core
from ..transport import Network


def _now_ms(t0: float) -> int:
    return int((time.monotonic() - t0) * 1000)


class BatchSchedulerService:
    """Serves ``/hypha-progress/0.0.1`` for one job.

    ``metrics`` receives ``(peer, Progress(kind='metrics'))``; ``done`` is set when every worker
    acknowledged the final update (batch_scheduler.rs:140-161)."""

    def __init__(self, network: Network, job_id: str, parameter_server: str, update_target: int,
                 update_rounds: int, time_cap_ms: int = 10000, update_cap: int = 3):
        self.core = load_core()
        self.network, self.job_id = network, job_id
        self._t0 = time.monotonic()
        tracker = self.core.ProgressTracker(parameter_server, update_target, update_rounds, 0)
        self.sched = self.core.BatchScheduler(tracker, time_cap_ms, update_cap)
        self.metrics: asyncio.Queue = asyncio.Queue(256)
        self.done = asyncio.Event()
        # collective jobs (enable_collective): `update` is a barrier over the CURRENT members; its answer
        # tells every rank who takes part in the outer step, so survivors never wait for a dead rank
        self._members: Optional[List[str]] = None
        self._slots: Dict[str, int] = {}       # member -> stable rank slot (a replacement takes a freed slot)
        self._generation = 0
        self._rendezvous_factory: Optional[Callable[[], str]] = None
        self._rendezvous: Optional[str] = None
        self._cond = asyncio.Condition()
        self._arrived: Dict[int, set] = {}     # round -> members whose `update` for that round arrived
        self._joiners: List[str] = []          # members that have not taken part in an outer step yet
        self._released: Dict[int, set] = {}    # round -> members that got the barrier's answer
        self._seen: set = set()                # workers that reported at least once
        self.never_started: List[str] = []     # late joiners dropped because the job ended before they trained
        self._lock = asyncio.Lock()
        self._reg = None
        self._kinds = {"status": self.core.ProgressKind.Status, "metrics": self.core.ProgressKind.Metrics,
                       "update": self.core.ProgressKind.Update, "updated": self.core.ProgressKind.Updated,
                       "update-received": self.core.ProgressKind.UpdateReceived}
        self._resp = {self.core.ResponseKind.Ok: "Ok", self.core.ResponseKind.Continue: "Continue",
                      self.core.ResponseKind.ScheduleUpdate: "ScheduleUpdate", self.core.ResponseKind.Done: "Done",
                      self.core.ResponseKind.Error: "Error"}

    def start_at_round(self, completed_rounds: int) -> None:
        """Resume: account for rounds that a previous run of the job already completed."""
        for _ in range(max(0, completed_rounds)):
            self.sched.tracker.next_round(_now_ms(self._t0))

    def add_worker(self, peer: str, batch_size: int) -> None:
        self.sched.tracker.worker_tracker.add_worker(peer, batch_size)

    def remove_worker(self, peer: str) -> None:
        self.sched.tracker.worker_tracker.remove_worker(peer)
        if self._members is not None and peer in self._members:
            self._members = [m for m in self._members if m != peer]
            self._slots.pop(peer, None)
            self._joiners = [j for j in self._joiners if j != peer]
            self._generation += 1
            self._rendezvous = self._rendezvous_factory() if self._rendezvous_factory else None
            asyncio.ensure_future(self._wake())       # the barrier may be complete without the lost rank

    def enable_collective(self, members: List[str], rendezvous: str, rendezvous_factory: Callable[[], str]) -> None:
        """Collective aggregation: remember the rank order and hand out a fresh rendezvous address every
        time the membership changes (the survivors re-form their process group on it)."""
        self._members, self._rendezvous, self._rendezvous_factory = list(members), rendezvous, rendezvous_factory
        self._slots = {m: i for i, m in enumerate(self._members)}

    def reserve_slot(self, peer: str) -> int:
        """The rank number a replacement will run under (lowest free one; needed BEFORE it is dispatched:
        engines that shard state by rank build their shard from it).  Kept by `add_member`."""
        if peer not in self._slots:
            taken = set(self._slots.values())
            self._slots[peer] = next(i for i in range(len(taken) + 1) if i not in taken)
        return self._slots[peer]

    def release_slot(self, peer: str) -> None:
        if self._members is None or peer not in self._members:
            self._slots.pop(peer, None)

    def add_member(self, peer: str) -> None:
        """Collective jobs: a replacement rank joins at the next `update` barrier (it is reported in
        `membership.joined` until it has taken part in one outer step)."""
        if self._members is None or peer in self._members:
            return
        self._members.append(peer)
        # engines whose state is sharded BY RANK (the fused GPU outer step) need the newcomer to take over a
        # dead rank's number: lowest free slot first, a new one only when the group grows
        self.reserve_slot(peer)
        self._joiners.append(peer)
        self._generation += 1
        self._rendezvous = self._rendezvous_factory() if self._rendezvous_factory else None
        asyncio.ensure_future(self._wake())

    async def _wake(self) -> None:
        async with self._cond:
            self._cond.notify_all()

    def _barrier_complete(self, rnd: int) -> bool:
        here = self._arrived.get(rnd, set())
        return all(m in here for m in (self._members or []))

    @property
    def round(self) -> int:
        return self.sched.tracker.round()

    async def handle(self, peer: str, progress: M.Progress) -> M.ProgressResponse:
        async with self._lock:
            if progress.kind == "metrics":
                if self.metrics.full():
                    self.metrics.get_nowait()
                self.metrics.put_nowait((peer, progress))
            if progress.kind in ("status", "update", "update-received"):
                self._seen.add(peer)
            r = self.sched.handle(peer, self._kinds[progress.kind], progress.batch_size, _now_ms(self._t0))
            all_done = r.all_done
            if (not all_done and progress.kind == "update-received"
                    and self._resp[r.kind] == "Done"):
                # a replacement that was dispatched too late to be admitted before the last round never
                # reports: it must not keep the finished job open
                wt = self.sched.tracker.worker_tracker
                for w in [w for w in wt.worker() if w not in self._seen]:
                    wt.remove_worker(w)
                    self.never_started.append(w)
                all_done = all(st == self.core.WorkerState.Done for st in wt.states())
            if all_done:
                self.done.set()
            resp = M.ProgressResponse(self._resp[r.kind], r.counter)
        if self._members is not None and progress.kind == "update" and resp.type == "Ok":
            # outside the lock: wait until every current member reached the sync point of THIS round (or was
            # dropped).  Arrivals are counted per round: a slow rank that is still acknowledging round r must
            # not satisfy a fast rank's barrier of round r+1.
            rnd = self.round
            self._arrived.setdefault(rnd, set()).add(peer)
            for old in [k for k in self._arrived if k < rnd - 1]:
                del self._arrived[old]
            async with self._cond:
                self._cond.notify_all()
                await self._cond.wait_for(lambda: self._barrier_complete(rnd))
            resp.membership = {"generation": self._generation, "members": list(self._members),
                               "rendezvous": self._rendezvous, "round": rnd,
                               "joined": [j for j in self._joiners if j in self._members],
                               "slots": {m: self._slots[m] for m in self._members if m in self._slots}}
            self._released.setdefault(rnd, set()).add(peer)
            if self._released[rnd] >= set(self._members):
                # everybody has the answer of this round's barrier: the newcomers are regular members now
                self._joiners = [j for j in self._joiners if j not in self._released[rnd]]
                for old in [k for k in self._released if k < rnd]:
                    del self._released[old]
        return resp

    def start(self) -> None:
        async def serve(peer, msg):
            req = M.ProgressRequest.from_wire(msg)
            return await self.handle(peer, req.progress)

        self._reg = self.network.on(M.PROGRESS_PROTOCOL,
                                    lambda m: isinstance(m, dict) and str(m.get("job_id")) == self.job_id, serve)

    def stop(self) -> None:
        if self._reg is not None:
            self._reg.unregister()
            self._reg = None


class DataSchedulerService:
    """Answers ``api::Request::Data{dataset}`` with ``(data_provider, slice index)``."""

    def __init__(self, network: Network, dataset: str, data_provider: str, num_slices: int):
        self.network, self.dataset, self.data_provider = network, dataset, data_provider
        self.tracker = load_core().SliceTracker(num_slices)
        self._reg = None

    def start(self) -> None:
        async def serve(peer, msg):
            ds = msg["Data"]["dataset"]
            if ds != self.dataset:
                return M.ApiResponse.data_not_found()
            return M.ApiResponse.data_success(self.data_provider, self.tracker.next(peer))

        self._reg = self.network.on(M.API_PROTOCOL, lambda m: isinstance(m, dict) and "Data" in m, serve)

    def remove_worker(self, peer: str) -> None:
        self.tracker.remove_worker(peer)

    def stop(self) -> None:
        if self._reg is not None:
            self._reg.unregister()
            self._reg = None


# -- metrics bridge -------------------------------------------------------------------------------
class Connector:
    async def send(self, worker_id: str, round: int, name: str, value: float) -> None:
        raise NotImplementedError


class NoOpConnector(Connector):
    async def send(self, worker_id, round, name, value) -> None:
        return None


class MemoryConnector(Connector):
    def __init__(self) -> None:
        self.records: List[Dict[str, Any]] = []

    async def send(self, worker_id, round, name, value) -> None:
        self.records.append({"worker_id": worker_id, "round": round, "metric_name": name, "value": value})


class JsonlConnector(Connector):
    def __init__(self, path: str):
        self.path = path

    async def send(self, worker_id, round, name, value) -> None:
        with open(self.path, "a") as f:
            f.write(json.dumps({"worker_id": worker_id, "round": round, "metric_name": name, "value": value}) + "\n")


class AimConnector(Connector):
    """POST one JSON ``AimMetrics`` per metric to ``http://<status_bridge>/status``
    (metrics_bridge.rs:18-24,126-145)."""

    def __init__(self, status_bridge: str):
        self.url = f"http://{status_bridge}/status"

    async def send(self, worker_id, round, name, value) -> None:
        import aiohttp

        async with aiohttp.ClientSession() as s:
            async with s.post(self.url, json={"worker_id": worker_id, "round": round, "metric_name": name,
                                              "value": value}) as r:
                r.raise_for_status()


class OtelConnector(Connector):
    """Training metrics as OpenTelemetry gauges ``hypha.job.<metric>`` (attributes: ``worker``, ``round``) on the
    scheduler's meter, next to the bandwidth counters: a deployment that ships OTLP to a collector
    (`telemetry_endpoint`) sees loss / throughput / step times per worker without running the AIM driver.  The
    reference forwards training metrics to AIM only (metrics_bridge.rs:126-145) and exports system metrics over OTLP
    (metrics.rs:33-75); this joins the two."""

    def __init__(self, telemetry: Any):
        self.meter = telemetry.meter
        self._gauges: Dict[str, Any] = {}

    async def send(self, worker_id, round, name, value) -> None:
        g = self._gauges.get(name)
        if g is None:
            safe = "".join(c if c.isalnum() or c in "._-" else "_" for c in str(name))[:63]
            g = self._gauges[name] = self.meter.create_gauge(f"hypha.job.{safe}")
        g.set(float(value), {"worker": str(worker_id), "round": int(round)})


class FanOutConnector(Connector):
    """Every metric to every sink; a sink that fails does not starve the others."""

    def __init__(self, *sinks: Connector):
        self.sinks = [s for s in sinks if s is not None]

    async def send(self, worker_id, round, name, value) -> None:
        for s in self.sinks:
            try:
                await s.send(worker_id, round, name, value)
            except Exception:      # noqa: BLE001 - a broken sink must not stop the others
                pass


class MetricsBridge:
    """Fan-in of ``(peer, Metrics)`` queues -> a connector."""

    def __init__(self, connector: Connector):
        self.connector = connector
        self._queues: List[asyncio.Queue] = []

    def register_stream(self, q: asyncio.Queue) -> None:
        self._queues.append(q)

    async def run(self) -> None:
        async def pump(q: asyncio.Queue):
            while True:
                peer, progress = await q.get()
                for name, value in progress.metrics.items():
                    try:
                        await self.connector.send(peer, progress.round, name, float(value))
                    except Exception:      # a broken sink must not stop training
                        pass

        await asyncio.gather(*(pump(q) for q in self._queues))
