"""Loopback end-to-end: gateway + data node + 2 train workers + 1 aggregate worker + scheduler in
one process over memory:// links, train executors as real subprocesses behind UDS bridges.
This is BASELINE config 1 (GPT-2 DiLoCo on CPU) at test scale; `bench/plumbing_gpt2.py` runs the
full GPT-2 small variant."""
import asyncio
import json
import os
import sys

import pytest

from hypha_b200 import config as C
from hypha_b200 import roles
from hypha_b200.executor.dataset import write_synthetic_token_slices
from hypha_b200.models.gpt2 import GPT2Config
from hypha_b200.scheduler import MemoryConnector


def _worker_cfg(gw, work_dir, executors, gpu=4.0):
    return C.builder(C.WorkerConfig).with_overrides({
        "listen_addresses": [], "gateway_addresses": [gw], "exclude_cidr": [], "work_dir": work_dir,
        "resources": {"gpu": gpu, "cpu": 4.0, "memory": 8.0, "storage": 10.0},
        "offer": {"price": 1.0, "floor": 0.5, "strategy": {"type": "whole"}}, "executors": executors}).build()


TRAIN_EXEC = {"class": "train", "name": "diloco-transformer", "runtime": "process", "cmd": "python",
              "args": ["-m", "hypha_b200.executor.training", "--socket", "{SOCKET_PATH}", "--work-dir", "{WORK_DIR}",
                       "--job", "{JOB_JSON}"]}
PS_EXEC = {"class": "aggregate", "name": "parameter-server", "runtime": "parameter-server"}


async def _run_job(tmp, model_cfg_writer, repo, job_overrides, dataset_writer, rounds=2):
    registry = os.path.join(tmp, "registry")
    os.makedirs(os.path.join(registry, repo))
    model_cfg_writer(os.path.join(registry, repo, "config.json"))
    os.environ["HYPHA_HF_LOCAL"] = registry
    os.environ["OMP_NUM_THREADS"] = "2"
    ds_dir = os.path.join(tmp, "toyset")
    dataset_writer(ds_dir)
    gw_addr = f"memory://gw-{os.getpid()}-{id(tmp)}"
    loop = asyncio.get_event_loop()
    started = [loop.create_future() for _ in range(5)]
    gw_cfg = C.builder(C.GatewayConfig).with_overrides({"listen_addresses": [gw_addr]}).build()
    bg = [asyncio.ensure_future(roles.run_gateway(gw_cfg, peer_id="gateway", started=started[0]))]
    await started[0]
    data_cfg = C.builder(C.DataConfig).with_overrides({"listen_addresses": [], "gateway_addresses": [gw_addr],
                                                       "exclude_cidr": [], "dataset_path": ds_dir}).build()
    bg.append(asyncio.ensure_future(roles.run_data(data_cfg, peer_id="data", started=started[1])))
    logs = []
    for i, (name, ex) in enumerate([("worker-a", [TRAIN_EXEC]), ("worker-b", [TRAIN_EXEC]), ("worker-ps", [PS_EXEC])]):
        wd = os.path.join(tmp, name)
        os.makedirs(wd)
        bg.append(asyncio.ensure_future(roles.run_worker(_worker_cfg(gw_addr, wd, ex), peer_id=name,
                                                         started=started[2 + i] if i < 3 else None, log=lambda m: (logs.append(m), print(m))[0])))
    nodes = [await f for f in started[1:]]
    over = {"listen_addresses": [], "gateway_addresses": [gw_addr], "exclude_cidr": [],
            "scheduler": {"job": job_overrides}}
    sched_cfg = C.builder(C.SchedulerConfig).with_overrides(over).build()
    sink = MemoryConnector()
    try:
        outcome = await asyncio.wait_for(
            roles.run_scheduler(sched_cfg, peer_id="scheduler", metrics_connector=sink, allocation_pause_s=0.7), 240)
    finally:
        for n in nodes + [started[0].result()]:
            n.stop.set()
        await asyncio.wait(bg, timeout=20)
    return outcome, sink, logs, nodes


def test_gpt2_diloco_two_workers_loopback(tmp_path):
    tmp = str(tmp_path)
    job = {
        "model": {"repository": "local/gpt2-tiny", "filenames": ["config.json"], "type": "causal-lm",
                  "input_names": ["input_ids", "labels"]},
        "preprocessor": None,
        "dataset": {"dataset": "toyset"},
        "rounds": {"avg_samples_between_updates": 40, "update_rounds": 2, "max_batch_size": 4},
        "inner_optimizer": {"learning-rate": 3e-3},
        "outer_optimizer": {"learning-rate": 0.7, "momentum": 0.9},
        "resources": {"num_workers": 2, "worker": {"gpu": 1.0, "cpu": 1.0, "memory": 1.0, "storage": 0.0},
                      "parameter_server": {"gpu": 0.0, "cpu": 1.0, "memory": 1.0, "storage": 0.0},
                      "worker_price": {"bid": 1.0, "max": 2.0}, "parameter_server_price": {"bid": 1.0, "max": 2.0}},
    }
    outcome, sink, logs, nodes = asyncio.run(_run_job(
        tmp, lambda p: GPT2Config.tiny().to_json(p), "local/gpt2-tiny", job,
        lambda d: write_synthetic_token_slices(d, num_slices=6, samples_per_slice=16, seq_len=32, vocab_size=512)))
    assert outcome.finished, logs[-20:]
    assert outcome.rounds == 2 and sorted(outcome.workers) == ["worker-a", "worker-b"]
    assert outcome.parameter_server == "worker-ps"
    losses = {}
    for r in sink.records:
        assert r["metric_name"] == "loss"
        losses.setdefault(r["worker_id"], {})[r["round"]] = r["value"]
    assert set(losses) == {"worker-a", "worker-b"} and all(set(v) == {1, 2} for v in losses.values())
    assert all(0 < v < 10 for w in losses.values() for v in w.values())
    assert any("Finished training of 2 DiLoCo update rounds" in l for l in logs)
    # bandwidth counters saw the tensor streams in both directions on the aggregate worker
    ps = next(n for n in nodes if n.peer_id == "worker-ps")
    snap = ps.network.counters.snapshot()
    assert snap["hypha.bandwidth.inbound.bytes"].get("/hypha-tensor-stream/push", 0) > 100_000
    assert snap["hypha.bandwidth.outbound.bytes"].get("/hypha-tensor-stream/push", 0) > 100_000
    # job work dirs are cleaned up
    for w in ("worker-a", "worker-b", "worker-ps"):
        assert [d for d in os.listdir(os.path.join(tmp, w)) if d.startswith("hypha-")] == []
