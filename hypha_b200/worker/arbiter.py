"""Worker arbiter: answers resource ads with priced, leased offers and guards job dispatch.

Reference: crates/worker/src/arbiter.rs -- constants :22-29, run loop :88-326, request
processing :328-437.  Filtering / scoring runs in the C++ core (``rank_requests``), the lease
table is the C++ ``ResourceLeaseManager``, the ad window is the C++ ``BatchWindow``
(100 messages or 200 ms).

State machine of a lease: tentative (500 ms, created when the offer is sent) -> accepted
(first ``RenewLease`` from the owning scheduler, 10 s, ``timeout = now + 10 s`` on every
renewal) -> expired (prune loop every 250 ms releases resources and cancels linked jobs).
"""
from __future__ import annotations

import asyncio
import time
from typing import Any, Callable, List

from .. import messages as M
from ..ops.native import load_core
from ..transport import Network, NetworkError
from .jobs import JobManager

WINDOW_LIMIT = 100
WINDOW_S = 0.2
OFFER_TIMEOUT_S = 0.5
PRUNE_INTERVAL_S = 0.25
LEASE_DURATION_S = 10.0


class Arbiter:
    def __init__(self, network: Network, capacity: M.Resources, job_manager: JobManager, *,
                 price: float = 100.0, floor: float = 100.0, strategy: str = "flexible", evaluator=None,
                 log: Callable[[str], None] = print):
        self.core = load_core()
        self.network, self.capacity, self.jobs = network, capacity, job_manager
        self.price, self.floor, self.strategy = price, floor, strategy
        self.evaluator = evaluator or self.core.WeightedResourceEvaluator()
        self.leases = self.core.ResourceLeaseManager(capacity.to_core())
        self.log = log
        self._regs: List[Any] = []
        self.offers_sent = 0
        job_manager.on_job_end = self._job_ended

    def _job_ended(self, job, status) -> None:
        """A FAILED job gives its lease back at once (the scheduler will stop renewing it anyway, but the
        10 s expiry would keep this worker from offering for a replacement job in the meantime).  Finished
        jobs keep the reference behaviour: the lease simply expires."""
        if status != M.JobStatus.FAILED:
            return
        others = [j for j in self.jobs.find_jobs_by_lease(job.lease_id)
                  if j.id != job.id and j.execution.status == M.JobStatus.RUNNING]
        if not others:
            try:
                self.leases.remove(job.lease_id)
                self.log(f"[arbiter] job {job.id[:8]} failed -> lease {job.lease_id[:8]} released")
            except KeyError:
                pass                     # already expired and pruned

    # -- handlers ---------------------------------------------------------------------------------
    async def _on_renew(self, peer: str, msg) -> M.ApiResponse:
        lease_id = str(msg["RenewLease"]["id"])
        try:
            lease = self.leases.get(lease_id)
        except KeyError:
            return M.ApiResponse.renew_failed()
        if lease.leasable.peer_id != peer:                 # only the owner may renew (arbiter.rs:156-185)
            return M.ApiResponse.renew_failed()
        renewed = self.leases.renew(lease_id, LEASE_DURATION_S)
        return M.ApiResponse.renewed(lease_id, renewed.timeout)

    async def _on_dispatch(self, peer: str, msg) -> M.ApiResponse:
        body = msg["DispatchJob"]
        try:
            lease = self.leases.get_by_peer(peer)
            if lease.is_expired():
                return M.ApiResponse.dispatch_failed()
            spec = M.JobSpec.from_wire(body["spec"])
            await self.jobs.execute(str(body["id"]), spec, lease.id, peer)
        except Exception as exc:
            self.log(f"[arbiter] dispatch from {peer} failed: {type(exc).__name__}: {exc}")
            return M.ApiResponse.dispatch_failed()
        return M.ApiResponse.dispatched(str(body["id"]), lease.timeout)

    # -- ad processing ---------------------------------------------------------------------------------
    async def process_requests(self, batch: List[tuple]) -> int:
        """``batch`` = [(origin_peer, RequestWorker)].  Returns the number of offers sent."""
        supported = set(self.jobs.supported())
        views = []
        for i, (peer, req) in enumerate(batch):
            ok = all(e in supported for e in req.spec.executor)
            views.append(self.core.WorkerRequestView(i, req.bid, req.spec.resources.to_core(), ok))
        ranked = self.core.rank_requests(views, self.capacity.to_core(), self.floor, self.evaluator)
        sent = 0
        for v in ranked:
            peer, req = batch[v.index]
            if req.timeout <= time.time():
                continue
            if self.strategy == "whole":
                resources, price = self.capacity, max(self.price, req.bid)
            else:
                resources, price = req.spec.resources, req.bid
            try:
                lease = self.leases.request(peer, resources.to_core(), OFFER_TIMEOUT_S)
            except RuntimeError:
                continue                                   # insufficient resources right now
            offer = M.ApiRequest.worker_offer(lease.id, req.id, price, resources, lease.timeout)
            asyncio.ensure_future(self._send_offer(peer, offer, lease.id))
            sent += 1
        self.offers_sent += sent
        return sent

    async def _send_offer(self, peer: str, offer: M.ApiRequest, lease_id: str) -> None:
        try:
            await self.network.request(peer, M.API_PROTOCOL, offer, timeout=5.0)
        except NetworkError:
            try:
                self.leases.remove(lease_id)
            except KeyError:
                pass

    async def prune_once(self) -> int:
        removed = self.leases.proceed()
        for lease in removed:
            for job in self.jobs.find_jobs_by_lease(lease.id):
                self.log(f"[arbiter] lease {lease.id} expired -> cancelling job {job.id}")
                job.execution.cancel()
        return len(removed)

    async def _on_membership(self, peer: str, msg) -> M.ApiResponse:
        """Scheduler -> aggregate job: the set of train workers still in the job (sent after a lease was
        lost).  Only the scheduler that owns the job may change it."""
        body = msg["JobMembership"]
        known = False
        for job in self.jobs.find_jobs_by_scheduler(peer):
            if job.spec.job_id == str(body["job_id"]):
                job.execution.set_members([str(w) for w in body["workers"]])
                known = True
        return M.ApiResponse.job_membership(known)

    # -- main loop -----------------------------------------------------------------------------------------
    async def run(self, stop: asyncio.Event) -> None:
        self._regs = [
            self.network.on(M.API_PROTOCOL, lambda m: isinstance(m, dict) and "RenewLease" in m, self._on_renew),
            self.network.on(M.API_PROTOCOL, lambda m: isinstance(m, dict) and "DispatchJob" in m, self._on_dispatch),
            self.network.on(M.API_PROTOCOL, lambda m: isinstance(m, dict) and "JobMembership" in m, self._on_membership),
        ]
        sub = await self.network.subscribe(M.WORKER_TOPIC)
        window = self.core.BatchWindow(WINDOW_LIMIT, WINDOW_S)

        async def prune_loop():
            while not stop.is_set():
                await asyncio.sleep(PRUNE_INTERVAL_S)
                try:
                    await self.prune_once()
                except Exception as exc:     # keep pruning
                    self.log(f"[arbiter] prune error: {exc}")

        pruner = asyncio.ensure_future(prune_loop())
        stopper = asyncio.ensure_future(stop.wait())
        try:
            while not stop.is_set():
                dl = window.next_deadline()
                timeout = None if dl < 0 else max(0.0, dl - time.monotonic())
                getter = asyncio.ensure_future(sub.get())
                done, _ = await asyncio.wait({getter, stopper}, timeout=timeout, return_when=asyncio.FIRST_COMPLETED)
                batch = None
                if getter in done:
                    try:
                        origin, data = getter.result()
                        req = M.RequestWorker.from_wire(M.decode_cbor(data))
                        batch = window.push((origin, req), time.monotonic())
                    except (NetworkError, M.MessageError, KeyError, ValueError, TypeError):
                        pass
                else:
                    getter.cancel()
                if batch is None:
                    batch = window.poll(time.monotonic())
                if batch:
                    await self.process_requests(list(batch))
        finally:
            pruner.cancel()
            stopper.cancel()
            for r in self._regs:
                r.unregister()
            await sub.unsubscribe()
