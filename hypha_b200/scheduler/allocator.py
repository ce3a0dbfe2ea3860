"""Scheduler-side allocation: gossip a request, collect priced offers, lease the winners.

Reference: crates/scheduler/src/allocator.rs (GreedyWorkerAllocator :24-167, aggregation
:169-418), crates/scheduler/src/worker.rs (lease-renewing Worker future :59-183),
crates/scheduler/src/task.rs (dispatch :27-121).

Offer ranking / diversity / moving deadline run in the C++ ``OfferAggregator``; this module is
the asyncio glue around it.
"""
from __future__ import annotations

import asyncio
import time
from dataclasses import dataclass
from typing import Callable, List, Optional

from .. import messages as M
from ..ops.native import load_core
from ..transport import Network, NetworkError

ALLOCATOR_DEADLINE_S = 5.0        # allocator.rs:25
RENEW_FALLBACK_S = 6.0            # worker.rs:103-117


class AllocatorError(Exception):
    pass


class NoOffersReceived(AllocatorError):
    pass


class LeaseExpired(Exception):
    pass


@dataclass
class Offer:
    peer_id: str
    lease_id: str
    request_id: str
    price: float
    resources: M.Resources
    timeout: float


class WorkerHandle:
    """A leased worker.  ``task`` resolves (raises) when the lease can no longer be renewed.

    Renewal cadence: immediately on creation (which turns the 500 ms tentative offer into a
    10 s lease, arbiter.rs:160), then after 2/3 of the remaining time (worker.rs:103-117)."""

    def __init__(self, network: Network, offer: Offer, clock: Callable[[], float] = time.time):
        self.network, self.offer = network, offer
        self.peer_id, self.lease_id = offer.peer_id, offer.lease_id
        self.resources, self.price = offer.resources, offer.price
        self.timeout = offer.timeout
        self._clock = clock
        self.task: Optional[asyncio.Task] = None
        self.renewals = 0

    @classmethod
    async def create(cls, network: Network, offer: Offer) -> "WorkerHandle":
        h = cls(network, offer)
        await h._renew_once()                    # accept the offer
        h.task = asyncio.ensure_future(h._renew_loop())
        return h

    async def _renew_once(self) -> None:
        try:
            resp = M.ApiResponse.from_wire(
                await self.network.request(self.peer_id, M.API_PROTOCOL, M.ApiRequest.renew_lease(self.lease_id),
                                           timeout=5.0))
        except NetworkError as exc:
            raise LeaseExpired(f"renew of {self.lease_id} on {self.peer_id} failed: {exc}") from exc
        kind, body = resp.variant()
        if resp.kind != "RenewLease" or kind != "Renewed":
            raise LeaseExpired(f"worker {self.peer_id} refused to renew {self.lease_id}")
        self.timeout = M.time_from_wire(body["timeout"])
        self.renewals += 1

    async def _renew_loop(self) -> None:
        while True:
            remaining = self.timeout - self._clock()
            delay = remaining * 2.0 / 3.0 if remaining > 0 else RENEW_FALLBACK_S
            await asyncio.sleep(max(0.0, delay))
            await self._renew_once()

    def failed(self) -> bool:
        return self.task is not None and self.task.done() and not self.task.cancelled()

    async def release(self) -> None:
        if self.task is not None:
            self.task.cancel()
            try:
                await self.task
            except (asyncio.CancelledError, Exception):
                pass


class GreedyWorkerAllocator:
    def __init__(self, network: Network, evaluator=None):
        self.network = network
        self.core = load_core()
        self.evaluator = evaluator or self.core.WeightedResourceEvaluator()

    async def request(self, spec: M.WorkerSpec, bid: float, max_price: float, num: int,
                      deadline_s: Optional[float] = None, diversity: bool = True) -> List[WorkerHandle]:
        """Advertise ``spec`` at ``bid`` on ``hypha/worker``; lease the best ``num`` offers."""
        deadline_s = ALLOCATOR_DEADLINE_S if deadline_s is None else deadline_s
        req_id = M.new_id()
        now = time.time()
        offers: asyncio.Queue = asyncio.Queue(100)

        async def on_offer(peer: str, msg):
            body = msg["WorkerOffer"]
            await offers.put(Offer(peer, str(body["id"]), str(body["request_id"]), float(body["price"]),
                                   M.Resources.from_wire(body["resources"]), M.time_from_wire(body["timeout"])))
            return M.ApiResponse.worker_offer()

        reg = self.network.on(M.API_PROTOCOL,
                              lambda m: isinstance(m, dict) and "WorkerOffer" in m
                              and str(m["WorkerOffer"].get("request_id")) == req_id, on_offer)
        agg = self.core.OfferAggregator(now=now, deadline_s=deadline_s, desired=num, diversity=diversity,
                                        upper_price=max_price, evaluator=self.evaluator)
        by_lease = {}
        try:
            ad = M.RequestWorker(req_id, spec, now + deadline_s, bid)
            await self.network.publish(M.WORKER_TOPIC, M.encode_cbor(ad))
            while not agg.done(time.time()):
                wait = max(0.0, agg.deadline() - time.time())
                try:
                    off = await asyncio.wait_for(offers.get(), timeout=wait if wait > 0 else 0.001)
                except asyncio.TimeoutError:
                    continue
                by_lease[off.lease_id] = off
                agg.push(self.core.Offer(off.peer_id, off.lease_id, off.request_id, off.price,
                                         off.resources.to_core(), off.timeout), time.time())
        finally:
            reg.unregister()
        chosen = [by_lease[o.lease_id] for o in agg.finish()]
        if not chosen:
            raise NoOffersReceived(f"no offers for request {req_id}")
        results = await asyncio.gather(*(WorkerHandle.create(self.network, o) for o in chosen),
                                       return_exceptions=True)
        handles = [r for r in results if isinstance(r, WorkerHandle)]
        if not handles:
            raise NoOffersReceived(f"every offer for request {req_id} expired before acceptance")
        return handles


class TaskError(Exception):
    pass


class Task:
    """A dispatched job on one or more workers; ``statuses`` streams ``JobStatus`` reports."""

    def __init__(self, task_id: str, spec: M.JobSpec, workers: List[WorkerHandle], network: Network):
        self.id, self.spec, self.workers, self.network = task_id, spec, workers, network
        self.statuses: asyncio.Queue = asyncio.Queue(16)
        self._reg = None

    @classmethod
    async def try_new(cls, network: Network, spec: M.JobSpec, workers: List[WorkerHandle]) -> "Task":
        task = cls(M.new_id(), spec, workers, network)

        async def on_status(peer, msg):
            await task.statuses.put((peer, msg["JobStatus"]["status"]["type"]))
            return M.ApiResponse.job_status()

        task._reg = network.on(M.API_PROTOCOL,
                               lambda m: isinstance(m, dict) and "JobStatus" in m
                               and str(m["JobStatus"].get("task_id")) == task.id, on_status)
        for w in workers:
            try:
                resp = M.ApiResponse.from_wire(
                    await network.request(w.peer_id, M.API_PROTOCOL, M.ApiRequest.dispatch_job(task.id, spec),
                                          timeout=30.0))
            except NetworkError as exc:
                task.close()
                raise TaskError(f"dispatch to {w.peer_id} failed: {exc}") from exc
            kind, _ = resp.variant()
            if resp.kind != "DispatchJob" or kind != "Dispatched":
                task.close()
                raise TaskError(f"worker {w.peer_id} rejected job {spec.job_id}")
        return task

    def close(self) -> None:
        if self._reg is not None:
            self._reg.unregister()
            self._reg = None
