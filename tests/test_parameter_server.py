"""Aggregate executor (`worker/jobs.py: ParameterServerExecutor`) against fake train workers: true 1/N mean
(the reference's pairwise running average is only right for N=2: SURVEY Q1), Nesterov numerics, the
`updated` report before the fan-out, receive-timeout membership shrink (Q8) and the scheduler-driven
membership changes with the cumulative update for a newcomer."""
import asyncio
import os

import torch
from safetensors.torch import load_file, save_file

from hypha_b200 import messages as M
from hypha_b200.transport.network import Network
from hypha_b200.worker.bridge import Connector
from hypha_b200.worker.jobs import ParameterServerExecutor


async def _cluster(tmp, workers, rounds, timeout_s=30.0):
    ps, sched = Network("ps"), Network("sched")
    addr = await ps.listen(f"memory://ps-test-{os.getpid()}-{id(ps)}")
    await sched.dial(addr)
    nets = {}
    for w in workers:
        nets[w] = Network(w)
        await nets[w].dial(addr)
    state = {"updated": 0, "order": []}

    async def progress(peer, msg):
        req = M.ProgressRequest.from_wire(msg)
        assert peer == "ps" and req.progress.kind == "updated"
        state["updated"] += 1
        state["order"].append(("updated", state["updated"]))
        return M.ProgressResponse("Done" if state["updated"] >= rounds else "Ok", 0)

    sched.on(M.PROGRESS_PROTOCOL, None, progress)
    logs = []
    ex = ParameterServerExecutor(Connector(ps), ps, tmp, receive_timeout_s=timeout_s, log=logs.append)
    return ps, sched, nets, state, logs, ex


def _spec(workers):
    cfg = M.AggregateExecutorConfig(M.Receive.from_peers(workers), M.Send.to_peers(workers), M.Nesterov(0.1, 0.7))
    return M.JobSpec("job-ps", M.Executor.aggregate(M.AGGREGATE_EXECUTOR_NAME, cfg))


async def _push(net, tmp, name, tensors):
    path = os.path.join(tmp, f"{name}.safetensors")
    save_file({k: v.contiguous() for k, v in tensors.items()}, path)
    await net.stream_push("ps", path, job_id="job-ps")


async def _result(q, tmp, tag):
    st = await asyncio.wait_for(q.get(), 10)
    path = os.path.join(tmp, f"result-{tag}")
    await st.save(path)
    return st.peer, load_file(path)


def test_true_mean_of_three_workers_and_reference_nesterov(tmp_path):
    tmp = str(tmp_path)

    async def main():
        workers = ["w1", "w2", "w3"]
        ps, sched, nets, state, logs, ex = await _cluster(tmp, workers, rounds=2)
        queues = {w: nets[w].streams_push("job-ps") for w in workers}
        run = await ex.execute(_spec(workers), "task-1", "sched")
        # round 0: pseudo-gradients 0.3, 0.6, 0.6 -> true mean 0.5 (pairwise running average would give 0.525)
        for w, g in zip(workers, (0.3, 0.6, 0.6)):
            await _push(nets[w], tmp, f"{w}-r0", {"w": torch.full((4,), g), "b": torch.full((2,), g)})
        for w in workers:
            peer, upd = await _result(queues[w], tmp, f"{w}-0")
            assert peer == "ps"
            assert torch.allclose(upd["w"], torch.full((4,), 0.085), atol=1e-7)      # lr=.1 mu=.7 g0=.5 -> .085
        assert state["updated"] == 1                            # reported BEFORE the results went out
        # round 1: g1 = [.1 .2 .3 .4] for everyone -> u1 = [.0415 .0585 .0755 .0925]
        g1 = torch.tensor([0.1, 0.2, 0.3, 0.4])
        for w in workers:
            await _push(nets[w], tmp, f"{w}-r1", {"w": g1, "b": torch.zeros(2)})
        for w in workers:
            _, upd = await _result(queues[w], tmp, f"{w}-1")
            assert torch.allclose(upd["w"], torch.tensor([0.0415, 0.0585, 0.0755, 0.0925]), atol=1e-6)
        assert await asyncio.wait_for(run.wait(), 10) == M.JobStatus.FINISHED
        assert state["updated"] == 2 and not os.path.exists(run.work_dir)       # work dir cleaned up
        for n in [ps, sched, *nets.values()]:
            await n.close()

    asyncio.run(main())


def test_receive_timeout_drops_the_silent_worker(tmp_path):
    tmp = str(tmp_path)

    async def main():
        workers = ["w1", "w2", "w3"]
        ps, sched, nets, state, logs, ex = await _cluster(tmp, workers, rounds=1, timeout_s=1.0)
        queues = {w: nets[w].streams_push("job-ps") for w in workers}
        run = await ex.execute(_spec(workers), "task-1", "sched")
        for w, g in (("w1", 0.2), ("w2", 0.4)):                 # w3 never sends
            await _push(nets[w], tmp, f"{w}-r0", {"w": torch.full((3,), g)})
        for w in ("w1", "w2"):
            _, upd = await _result(queues[w], tmp, w)
            assert torch.allclose(upd["w"], torch.full((3,), 0.1 * (0.7 * 0.3 + 0.3)), atol=1e-7)   # mean of the two
        assert queues["w3"].empty()
        assert await asyncio.wait_for(run.wait(), 10) == M.JobStatus.FINISHED
        assert any("timeout, dropping ['w3']" in l for l in logs)
        for n in [ps, sched, *nets.values()]:
            await n.close()

    asyncio.run(main())


def test_membership_changes_and_cumulative_update_for_a_newcomer(tmp_path):
    tmp = str(tmp_path)

    async def main():
        workers = ["w1", "w2", "w3"]
        ps, sched, nets, state, logs, ex = await _cluster(tmp, workers, rounds=3)
        nets["w4"] = Network("w4")
        await nets["w4"].dial(ps.listen_addrs[0])
        queues = {w: nets[w].streams_push("job-ps") for w in ("w1", "w2", "w3", "w4")}
        run = await ex.execute(_spec(workers), "task-1", "sched")
        # round 0 with w1, w2 only: the scheduler removes w3 while the PS is waiting for its file
        for w in ("w1", "w2"):
            await _push(nets[w], tmp, f"{w}-r0", {"w": torch.full((2,), 1.0)})
        await asyncio.sleep(0.2)
        run.set_members(["w1", "w2"])
        u0 = (await _result(queues["w1"], tmp, "w1-0"))[1]["w"]
        await _result(queues["w2"], tmp, "w2-0")
        assert torch.allclose(u0, torch.full((2,), 0.1 * (0.7 * 1.0 + 1.0)))
        # a replacement is announced mid-round 1: admitted at the next boundary with U = u0 + u1
        await _push(nets["w1"], tmp, "w1-r1", {"w": torch.full((2,), 0.5)})
        await asyncio.sleep(0.2)
        run.set_members(["w1", "w2", "w4"])
        await asyncio.sleep(0.2)
        await _push(nets["w4"], tmp, "w4-early", {"w": torch.full((2,), 9.0)})      # not admitted yet: ignored
        await _push(nets["w2"], tmp, "w2-r1", {"w": torch.full((2,), 0.5)})
        u1 = (await _result(queues["w1"], tmp, "w1-1"))[1]["w"]
        await _result(queues["w2"], tmp, "w2-1")
        peer, cum = await _result(queues["w4"], tmp, "w4-cum")
        assert peer == "ps" and torch.allclose(cum["w"], u0 + u1, atol=1e-7)
        for _ in range(100):                  # the PS logs after its send returned; the receiver may be quicker
            if any("w4 joined (cumulative update sent)" in l for l in logs):
                break
            await asyncio.sleep(0.02)
        assert any("w4 joined (cumulative update sent)" in l for l in logs)
        # round 2: all three current members contribute and all three get the result
        for w in ("w1", "w2", "w4"):
            await _push(nets[w], tmp, f"{w}-r2", {"w": torch.full((2,), 0.3)})
        for w in ("w1", "w2", "w4"):
            await _result(queues[w], tmp, f"{w}-2")
        assert await asyncio.wait_for(run.wait(), 10) == M.JobStatus.FINISHED
        for n in [ps, sched, *nets.values()]:
            await n.close()

    asyncio.run(main())
