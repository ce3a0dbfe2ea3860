"""Allocation protocol end to end over in-memory links: gossip ad -> priced offers -> lease
acceptance by renewal -> dispatch guard -> expiry.  (The reference has no tests for the arbiter /
allocator / lease manager: SURVEY section 4 "not tested at all".)"""
import asyncio
import json
import os
import time

import pytest

from hypha_b200 import messages as M
from hypha_b200.scheduler import GreedyWorkerAllocator, NoOffersReceived
from hypha_b200.transport import Network
from hypha_b200.worker import Arbiter, Connector, JobManager
from hypha_b200.worker import arbiter as arb_mod


class _Exec:
    def __init__(self, cls, name, runtime="parameter-server"):
        self.cls, self.name, self.runtime, self.cmd, self.args = cls, name, runtime, None, []


async def _cluster(worker_specs):
    """worker_specs: [(peer_id, gpu, price, floor, strategy)] -> (gateway, scheduler net, workers)"""
    gw = Network("gw", relay_server=True, dht_server=True)
    addr = await gw.listen(f"memory://alloc-{id(gw)}")
    sched = Network("sched")
    await sched.dial(addr)
    workers = []
    for pid, gpu, price, floor, strategy in worker_specs:
        net = Network(pid)
        await net.listen(f"memory://{pid}-{id(net)}")
        await net.dial(addr)
        jobs = JobManager(Connector(net), net, "/tmp", [_Exec("train", M.TRAIN_EXECUTOR_NAME), _Exec("aggregate", M.AGGREGATE_EXECUTOR_NAME)],
                          log=lambda m: None)
        a = Arbiter(net, M.Resources(gpu=gpu, cpu=8, storage=100, memory=64), jobs, price=price, floor=floor,
                    strategy=strategy, log=lambda m: None)
        stop = asyncio.Event()
        task = asyncio.ensure_future(a.run(stop))
        workers.append((net, a, stop, task))
    await asyncio.sleep(0.05)          # subscriptions propagate
    return gw, sched, workers


async def _teardown(gw, sched, workers):
    for net, a, stop, task in workers:
        stop.set()
    await asyncio.gather(*(w[3] for w in workers), return_exceptions=True)
    for n in [gw, sched] + [w[0] for w in workers]:
        await n.close()


SPEC = M.WorkerSpec(M.Resources(gpu=1, cpu=1, storage=0, memory=1), [M.ExecutorDescriptor("train", M.TRAIN_EXECUTOR_NAME)])


def test_cheapest_offers_win_with_diversity_and_leases_are_renewed():
    async def main():
        gw, sched, workers = await _cluster([("w-cheap", 4, 10.0, 1.0, "whole"), ("w-mid", 4, 20.0, 1.0, "whole"),
                                             ("w-dear", 4, 90.0, 1.0, "whole")])
        alloc = GreedyWorkerAllocator(sched)
        handles = await alloc.request(SPEC, bid=5.0, max_price=50.0, num=2, deadline_s=2.0)
        assert sorted(h.peer_id for h in handles) == ["w-cheap", "w-mid"]      # 90 > max price, best two kept
        assert all(h.renewals >= 1 for h in handles)                            # accepted by renewing
        for h in handles:
            assert h.timeout - time.time() > 8.0                                # 10 s lease after acceptance
            assert h.resources.gpu == 4                                         # "whole" strategy offers everything
        by = {w[0].peer_id: w[1] for w in workers}
        assert by["w-cheap"].leases.available().gpu == 0
        await asyncio.sleep(arb_mod.OFFER_TIMEOUT_S + 2 * arb_mod.PRUNE_INTERVAL_S + 0.2)
        assert by["w-dear"].leases.available().gpu == 4                         # loser's tentative lease expired
        assert by["w-cheap"].leases.available().gpu == 0                        # winner's lease still held
        for h in handles:
            await h.release()
        await _teardown(gw, sched, workers)
    asyncio.run(main())


def test_floor_unsupported_executor_and_capacity_filters():
    async def main():
        gw, sched, workers = await _cluster([("w1", 2, 10.0, 50.0, "flexible")])
        alloc = GreedyWorkerAllocator(sched)
        with pytest.raises(NoOffersReceived):                                   # bid below the worker's floor
            await alloc.request(SPEC, bid=5.0, max_price=100.0, num=1, deadline_s=0.6)
        bad = M.WorkerSpec(SPEC.resources, [M.ExecutorDescriptor("train", "unknown-executor")])
        with pytest.raises(NoOffersReceived):
            await alloc.request(bad, bid=60.0, max_price=100.0, num=1, deadline_s=0.6)
        big = M.WorkerSpec(M.Resources(gpu=64, cpu=64, storage=1000, memory=1000), SPEC.executor)
        with pytest.raises(NoOffersReceived):                                   # strictly greater than capacity
            await alloc.request(big, bid=60.0, max_price=100.0, num=1, deadline_s=0.6)
        ok = await alloc.request(SPEC, bid=60.0, max_price=100.0, num=1, deadline_s=2.0)
        assert ok[0].peer_id == "w1" and ok[0].price == 60.0 and ok[0].resources.gpu == 1    # flexible mirrors the ask
        await ok[0].release()
        await _teardown(gw, sched, workers)
    asyncio.run(main())


def test_renew_by_non_owner_fails_and_dispatch_needs_a_lease():
    async def main():
        gw, sched, workers = await _cluster([("w1", 2, 10.0, 1.0, "flexible")])
        intruder = Network("intruder")
        await intruder.dial(gw.listen_addrs[0])
        alloc = GreedyWorkerAllocator(sched)
        (h,) = await alloc.request(SPEC, bid=20.0, max_price=100.0, num=1, deadline_s=2.0)
        resp = M.ApiResponse.from_wire(await intruder.request("w1", M.API_PROTOCOL, M.ApiRequest.renew_lease(h.lease_id)))
        assert resp.variant()[0] == "Failed"
        spec = M.JobSpec(M.new_id(), M.Executor.aggregate(M.AGGREGATE_EXECUTOR_NAME, M.AggregateExecutorConfig(
            M.Receive.from_peers(["x"]), M.Send.to_peers(["x"]), M.Nesterov(0.7, 0.9))))
        resp = M.ApiResponse.from_wire(await intruder.request("w1", M.API_PROTOCOL, M.ApiRequest.dispatch_job(M.new_id(), spec)))
        assert resp.variant()[0] == "Failed"                                    # no lease for this peer
        resp = M.ApiResponse.from_wire(await sched.request("w1", M.API_PROTOCOL, M.ApiRequest.renew_lease("no-such-lease")))
        assert resp.variant()[0] == "Failed"
        await h.release()
        await intruder.close()
        await _teardown(gw, sched, workers)
    asyncio.run(main())


def test_lease_expiry_cancels_the_job_and_fails_the_handle(monkeypatch):
    async def main():
        monkeypatch.setattr(arb_mod, "LEASE_DURATION_S", 0.6)
        gw, sched, workers = await _cluster([("w1", 2, 10.0, 1.0, "flexible")])
        net, a, stop, task = workers[0]
        alloc = GreedyWorkerAllocator(sched)
        aspec = M.WorkerSpec(SPEC.resources, [M.ExecutorDescriptor("aggregate", M.AGGREGATE_EXECUTOR_NAME)])
        (h,) = await alloc.request(aspec, bid=20.0, max_price=100.0, num=1, deadline_s=2.0)
        spec = M.JobSpec(M.new_id(), M.Executor.aggregate(M.AGGREGATE_EXECUTOR_NAME, M.AggregateExecutorConfig(
            M.Receive.from_peers(["nobody"]), M.Send.to_peers(["nobody"]), M.Nesterov(0.7, 0.9))))
        tid = M.new_id()
        resp = M.ApiResponse.from_wire(await sched.request("w1", M.API_PROTOCOL, M.ApiRequest.dispatch_job(tid, spec)))
        assert resp.variant()[0] == "Dispatched"
        job = a.jobs.jobs[tid]
        await h.release()                               # scheduler stops renewing -> lease runs out
        await asyncio.wait_for(job.execution.wait(), 5.0)
        assert job.execution.status == M.JobStatus.FAILED and a.leases.available().gpu == 2
        await _teardown(gw, sched, workers)
    asyncio.run(main())


def test_scheduler_without_workers_fails_with_a_role_error():
    """Nobody answers the advertisement: `run_scheduler` raises RoleError("allocation failed: ...") after the
    5 s allocator deadline (the CLI turns that into exit code 1 with a one-line message, no traceback)."""
    import tempfile
    import time

    from hypha_b200 import config as C
    from hypha_b200 import roles
    from hypha_b200.executor.dataset import write_synthetic_token_slices

    async def main():
        tmp = tempfile.mkdtemp()
        ds = os.path.join(tmp, "toyset")
        write_synthetic_token_slices(ds, 2, 4, 8, 32)
        gw = f"memory://gw-noworkers-{os.getpid()}"
        loop = asyncio.get_event_loop()
        s0, s1 = loop.create_future(), loop.create_future()
        bg = [asyncio.ensure_future(roles.run_gateway(
            C.builder(C.GatewayConfig).with_overrides({"listen_addresses": [gw]}).build(), peer_id="gw", started=s0))]
        await s0
        bg.append(asyncio.ensure_future(roles.run_data(C.builder(C.DataConfig).with_overrides(
            {"listen_addresses": [], "gateway_addresses": [gw], "exclude_cidr": [], "dataset_path": ds}).build(),
            peer_id="data", started=s1)))
        await s1
        cfg = C.builder(C.SchedulerConfig).with_overrides(
            {"listen_addresses": [], "gateway_addresses": [gw], "exclude_cidr": [],
             "scheduler": {"job": {"dataset": {"dataset": "toyset"}, "preprocessor": None}}}).build()
        t0 = time.time()
        with pytest.raises(roles.RoleError) as ei:
            await roles.run_scheduler(cfg, peer_id="sched")
        assert "allocation failed" in str(ei.value) and 4.0 < time.time() - t0 < 15.0
        for n in (s0.result(), s1.result()):
            n.stop.set()
        await asyncio.wait(bg, timeout=5)

    asyncio.run(main())


def test_collective_membership_barrier_semantics():
    """`update` of a collective job is a barrier over the current members of the current round: it is
    released by the last arrival or by the removal of the missing rank, its answer names the membership
    generation, and a newcomer is reported in `joined` for exactly one barrier."""
    from hypha_b200.scheduler.scheduling import BatchSchedulerService
    from hypha_b200.transport import Network

    async def main():
        svc = BatchSchedulerService(Network("sched-unit"), "job", "a", update_target=12, update_rounds=3)
        for w in ("a", "b", "c"):
            svc.add_worker(w, 4)
        ports = iter(range(40000, 40100))
        svc.enable_collective(["a", "b", "c"], "tcp://127.0.0.1:1", lambda: f"tcp://127.0.0.1:{next(ports)}")
        upd = M.Progress("update")

        # round 0: a and b wait until c is removed
        ta = asyncio.ensure_future(svc.handle("a", upd))
        tb = asyncio.ensure_future(svc.handle("b", upd))
        await asyncio.sleep(0.1)
        assert not ta.done() and not tb.done()
        svc.remove_worker("c")
        ra, rb = await asyncio.wait_for(asyncio.gather(ta, tb), 2)
        for r in (ra, rb):
            assert r.type == "Ok" and r.membership["generation"] == 1 and r.membership["members"] == ["a", "b"]
            assert r.membership["joined"] == [] and r.membership["round"] == 0
            assert r.membership["rendezvous"] == "tcp://127.0.0.1:40000"
            assert r.membership["slots"] == {"a": 0, "b": 1}                   # c's slot 2 is free now
        assert (await svc.handle("a", M.Progress("updated"))).type == "Ok"
        assert (await svc.handle("b", M.Progress("updated"))).type == "Ok"          # idempotent
        assert svc.round == 1
        for w in ("a", "b"):
            assert (await svc.handle(w, M.Progress("update-received"))).type == "Continue"

        # round 1: a newcomer d is admitted; a fast rank's barrier must wait for everybody of THIS round
        svc.add_worker("d", 4)
        svc.add_member("d")
        td = asyncio.ensure_future(svc.handle("d", upd))
        ta = asyncio.ensure_future(svc.handle("a", upd))
        await asyncio.sleep(0.1)
        assert not td.done() and not ta.done()
        rb = await asyncio.wait_for(svc.handle("b", upd), 2)
        ra, rd = await asyncio.wait_for(asyncio.gather(ta, td), 2)
        for r in (ra, rb, rd):
            assert r.membership["generation"] == 2 and r.membership["members"] == ["a", "b", "d"]
            assert r.membership["joined"] == ["d"] and r.membership["round"] == 1
            # stable rank numbering for engines that shard state by rank: d took over the dead rank's slot
            assert r.membership["slots"] == {"a": 0, "b": 1, "d": 2}
        await svc.handle("a", M.Progress("updated"))
        for w in ("a", "b", "d"):
            await svc.handle(w, M.Progress("update-received"))

        # round 2: d is a regular member now
        tasks = [asyncio.ensure_future(svc.handle(w, upd)) for w in ("a", "b", "d")]
        rs = await asyncio.wait_for(asyncio.gather(*tasks), 2)
        assert all(r.membership["joined"] == [] and r.membership["generation"] == 2 for r in rs)

    asyncio.run(main())


def test_allocation_scales_to_hundreds_of_workers():
    """256 arbiters behind one gateway answer one ad; the allocator collects one offer per peer, accepts all of
    them by renewal and returns early -- in well under the 5 s deadline (the 200 ms ad window dominates)."""
    async def main():
        n = 256
        gw, sched, workers = await _cluster([(f"w{i:03d}", 8, 10.0 + i % 7, 1.0, "flexible") for i in range(n)])
        alloc = GreedyWorkerAllocator(sched)
        t0 = time.perf_counter()
        handles = await alloc.request(SPEC, bid=50.0, max_price=100.0, num=n, deadline_s=5.0)
        dt = time.perf_counter() - t0
        assert len({h.peer_id for h in handles}) == n and dt < 3.0, dt
        assert all(h.renewals >= 1 for h in handles)
        for h in handles:
            await h.release()
        await _teardown(gw, sched, workers)
    asyncio.run(main())


def test_roles_survive_a_gateway_restart(monkeypatch):
    """The gateway process is replaced (same address, empty DHT): the data node re-dials it, announces its
    dataset again and a scheduler that arrives later finds the dataset and the worker's offers through the new
    gateway."""
    import tempfile

    from hypha_b200 import config as C
    from hypha_b200 import roles
    from hypha_b200.executor.dataset import write_synthetic_token_slices
    from hypha_b200.transport import Network, RecordNotFound

    monkeypatch.setattr(roles, "GATEWAY_REDIAL_S", 0.2)

    async def main():
        tmp = tempfile.mkdtemp()
        ds = os.path.join(tmp, "toyset")
        write_synthetic_token_slices(ds, 2, 4, 8, 32)
        gw = f"memory://gw-restart-{os.getpid()}"
        loop = asyncio.get_event_loop()

        async def start_gateway(pid):
            fut = loop.create_future()
            t = asyncio.ensure_future(roles.run_gateway(
                C.builder(C.GatewayConfig).with_overrides({"listen_addresses": [gw]}).build(), peer_id=pid, started=fut))
            return t, await fut

        t1, g1 = await start_gateway("gw-1")
        s1 = loop.create_future()
        data_task = asyncio.ensure_future(roles.run_data(C.builder(C.DataConfig).with_overrides(
            {"listen_addresses": [], "gateway_addresses": [gw], "exclude_cidr": [], "dataset_path": ds}).build(),
            peer_id="data", started=s1))
        data = await s1
        assert data.gateways == {gw: "gw-1"}
        g1.stop.set()                                   # the gateway goes away ...
        await asyncio.wait([t1], timeout=5)
        await asyncio.sleep(0.5)
        assert "gw-1" not in data.network.links          # noticed and dropped by the maintenance loop
        t2, g2 = await start_gateway("gw-2")            # ... and comes back as a new process
        for _ in range(100):
            if data.gateways.get(gw) == "gw-2" and "gw-2" in data.network.links:
                break
            await asyncio.sleep(0.05)
        assert data.gateways == {gw: "gw-2"}
        probe = Network("late-scheduler")
        await probe.dial(gw)
        rec = None
        for _ in range(100):
            try:
                rec = await probe.get_record("toyset")
                break
            except RecordNotFound:
                await asyncio.sleep(0.05)
        assert rec is not None and json.loads(rec[0])["num_slices"] == 2 and rec[1] == "data"   # re-announced into the new DHT
        await probe.close()
        for n in (data, g2):
            n.stop.set()
        await asyncio.wait([data_task, t2], timeout=5)

    asyncio.run(main())
