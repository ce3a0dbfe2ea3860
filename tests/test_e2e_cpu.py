"""Loopback end-to-end: gateway + data node + 2 train workers + 1 aggregate worker + scheduler in
one process over memory:// links, train executors as real subprocesses behind UDS bridges.
This is BASELINE config 1 (GPT-2 DiLoCo on CPU) at test scale; `bench/plumbing_gpt2.py` runs the
full GPT-2 small variant."""
import asyncio
import json
import os

import pytest
import torch

from hypha_b200.executor.dataset import write_synthetic_token_slices
from hypha_b200.models.gpt2 import GPT2Config
from hypha_b200.local_cluster import run_loopback_job


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
    outcome, sink, logs, nodes = asyncio.run(run_loopback_job(
        tmp, repo="local/gpt2-tiny", write_model_config=lambda p: GPT2Config.tiny().to_json(p), job=job,
        write_dataset=lambda d: write_synthetic_token_slices(d, num_slices=6, samples_per_slice=16, seq_len=32,
                                                             vocab_size=512), timeout_s=240))
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


def test_gpt2_diloco_collective_mode_loopback(tmp_path):
    """`aggregation.mode = "collective"`: no aggregate worker is allocated; the two train executors
    form a gloo process group over the rendezvous the scheduler hands out, all-reduce the
    pseudo-gradient, apply Nesterov locally and rank 0 reports `updated`.  On a GPU box the same job
    runs the fused peer-memory outer step (tests/test_multigpu.py)."""
    tmp = str(tmp_path)
    job = {
        "model": {"repository": "local/gpt2-tiny", "filenames": ["config.json"], "type": "causal-lm",
                  "input_names": ["input_ids", "labels"]},
        "preprocessor": None,
        "dataset": {"dataset": "toyset"},
        "rounds": {"avg_samples_between_updates": 40, "update_rounds": 2, "max_batch_size": 4},
        "inner_optimizer": {"learning-rate": 3e-3},
        "outer_optimizer": {"learning-rate": 0.7, "momentum": 0.9},
        "aggregation": {"mode": "collective"},
        "resources": {"num_workers": 2, "worker": {"gpu": 1.0, "cpu": 1.0, "memory": 1.0, "storage": 0.0},
                      "parameter_server": {"gpu": 0.0, "cpu": 1.0, "memory": 1.0, "storage": 0.0},
                      "worker_price": {"bid": 1.0, "max": 2.0}, "parameter_server_price": {"bid": 1.0, "max": 2.0}},
    }
    os.environ["HYPHA_EXECUTOR_DEVICE"] = "cpu"
    try:
        outcome, sink, logs, nodes = asyncio.run(run_loopback_job(
            tmp, repo="local/gpt2-tiny", write_model_config=lambda p: GPT2Config.tiny().to_json(p), job=job,
            write_dataset=lambda d: write_synthetic_token_slices(d, num_slices=6, samples_per_slice=16, seq_len=32,
                                                                 vocab_size=512), timeout_s=240))
    finally:
        os.environ.pop("HYPHA_EXECUTOR_DEVICE", None)
    assert outcome.finished, logs[-20:]
    assert outcome.rounds == 2 and sorted(outcome.workers) == ["worker-a", "worker-b"]
    assert outcome.parameter_server == "collective"
    assert sum("outer step via all-reduce" in l for l in logs) == 2
    names = {r["metric_name"] for r in sink.records}
    assert names == {"loss", "samples_or_tokens_per_s", "inner_step_ms", "outer_step_ms", "inner_steps", "tokens"}
    assert all(r["value"] > 0 for r in sink.records)
    losses = {}
    for r in sink.records:
        if r["metric_name"] != "loss":
            continue
        losses.setdefault(r["worker_id"], {})[r["round"]] = r["value"]
    assert set(losses) == {"worker-a", "worker-b"} and all(set(v) == {1, 2} for v in losses.values())
    assert sum("Finished training of 2 DiLoCo update rounds" in l for l in logs) == 2


def test_scheduler_survives_a_lost_train_worker(tmp_path):
    """Elastic job (quirks Q8/Q9): one of three train workers dies after the first round.  The scheduler
    notices the lost lease, removes the worker from the batch and slice schedulers, tells the aggregate
    job the new membership (so it stops waiting for that worker's pseudo-gradient) and the job
    finishes all rounds with the two survivors.  The reference scheduler exits at this point."""
    tmp = str(tmp_path)
    job = {
        "model": {"repository": "local/gpt2-tiny", "filenames": ["config.json"], "type": "causal-lm",
                  "input_names": ["input_ids", "labels"]},
        "preprocessor": None,
        "dataset": {"dataset": "toyset"},
        "rounds": {"avg_samples_between_updates": 48, "update_rounds": 3, "max_batch_size": 4},
        "inner_optimizer": {"learning-rate": 3e-3},
        "outer_optimizer": {"learning-rate": 0.7, "momentum": 0.9},
        "resources": {"num_workers": 3, "worker": {"gpu": 1.0, "cpu": 1.0, "memory": 1.0, "storage": 0.0},
                      "parameter_server": {"gpu": 0.0, "cpu": 1.0, "memory": 1.0, "storage": 0.0},
                      "worker_price": {"bid": 1.0, "max": 2.0}, "parameter_server_price": {"bid": 1.0, "max": 2.0}},
    }

    async def kill_worker_c(nodes, logs):
        while not any(l.startswith("round 1:") or " round 1:" in l for l in logs):
            await asyncio.sleep(0.1)
        nodes["worker-c"].stop.set()

    outcome, sink, logs, nodes = asyncio.run(run_loopback_job(
        tmp, repo="local/gpt2-tiny", write_model_config=lambda p: GPT2Config.tiny().to_json(p), job=job, num_workers=3,
        write_dataset=lambda d: write_synthetic_token_slices(d, num_slices=9, samples_per_slice=16, seq_len=32,
                                                             vocab_size=512), timeout_s=300, chaos=kill_worker_c))
    assert outcome.finished, logs[-25:]
    assert outcome.rounds == 3 and outcome.removed_workers == ["worker-c"]
    assert any("scheduler removed [\'worker-c\']" in l or "dropping [\'worker-c\']" in l for l in logs), logs[-25:]
    done = [l for l in logs if "Finished training of 3 DiLoCo update rounds" in l]
    assert len(done) == 2                                   # the two survivors ran to completion
    last_round = {}
    for r in sink.records:
        last_round[r["worker_id"]] = max(last_round.get(r["worker_id"], 0), r["round"])
    assert last_round["worker-a"] == 3 and last_round["worker-b"] == 3


def test_lost_worker_is_replaced_and_the_newcomer_joins(tmp_path):
    """Rejoin through the control plane (quirk Q9, parameter-server mode): three train workers plus one
    more expensive spare.  worker-c dies after round 1; the scheduler drops it, leases the spare,
    dispatches a `join` train job and tells the aggregate worker the new membership; at the next round
    boundary the aggregate worker sends the newcomer the cumulative update U, so it starts from
    theta_0(seed) + U = the current global model, and takes part in the remaining rounds."""
    tmp = str(tmp_path)
    job = {
        "model": {"repository": "local/gpt2-tiny", "filenames": ["config.json"], "type": "causal-lm",
                  "input_names": ["input_ids", "labels"]},
        "preprocessor": None,
        "dataset": {"dataset": "toyset"},
        "rounds": {"avg_samples_between_updates": 48, "update_rounds": 10, "max_batch_size": 4},
        "inner_optimizer": {"learning-rate": 3e-3},
        "outer_optimizer": {"learning-rate": 0.7, "momentum": 0.9},
        "resources": {"num_workers": 3, "worker": {"gpu": 1.0, "cpu": 1.0, "memory": 1.0, "storage": 0.0},
                      "parameter_server": {"gpu": 0.0, "cpu": 1.0, "memory": 1.0, "storage": 0.0},
                      "worker_price": {"bid": 1.0, "max": 2.0}, "parameter_server_price": {"bid": 1.0, "max": 2.0}},
    }

    async def kill_worker_c(nodes, logs):
        while not any(" round 1:" in l or l.startswith("round 1:") for l in logs):
            await asyncio.sleep(0.1)
        nodes["worker-c"].stop.set()

    outcome, sink, logs, nodes = asyncio.run(run_loopback_job(
        tmp, repo="local/gpt2-tiny", write_model_config=lambda p: GPT2Config.tiny().to_json(p), job=job, num_workers=3,
        spare_workers=1, chaos=kill_worker_c, timeout_s=400,
        write_dataset=lambda d: write_synthetic_token_slices(d, num_slices=12, samples_per_slice=16, seq_len=32,
                                                             vocab_size=512)))
    assert outcome.finished, logs[-30:]
    assert outcome.rounds == 10
    assert sorted(outcome.workers) == ["worker-a", "worker-b", "worker-c"]       # the initial allocation
    assert outcome.removed_workers == ["worker-c"] and outcome.joined_workers == ["worker-d"]
    assert any("worker-d joined (cumulative update sent)" in l for l in logs), logs[-30:]
    assert any("joined a running job after round" in l for l in logs)
    rounds = {}
    for r in sink.records:
        rounds.setdefault(r["worker_id"], set()).add(r["round"])
    assert rounds["worker-a"] == set(range(1, 11)) and rounds["worker-b"] == set(range(1, 11))
    assert len(rounds.get("worker-d", set())) >= 2                                  # it trained for real
    assert sum("Finished training of" in l for l in logs) == 3                      # a, b and the newcomer


def test_crashed_executor_is_restarted_on_the_same_worker(tmp_path):
    """The worker role stays up but its executor process is killed: the worker reports JobStatus Failed,
    the scheduler drops the rank, frees the lease, and -- the worker being the cheapest offer again --
    re-leases it and dispatches a `join` job, which resumes from the cumulative update."""
    import signal

    tmp = str(tmp_path)
    job = {
        "model": {"repository": "local/gpt2-tiny", "filenames": ["config.json"], "type": "causal-lm",
                  "input_names": ["input_ids", "labels"]},
        "preprocessor": None,
        "dataset": {"dataset": "toyset"},
        "rounds": {"avg_samples_between_updates": 480, "update_rounds": 30, "max_batch_size": 4},
        "inner_optimizer": {"learning-rate": 3e-3},
        "outer_optimizer": {"learning-rate": 0.7, "momentum": 0.9},
        "resources": {"num_workers": 3, "worker": {"gpu": 1.0, "cpu": 1.0, "memory": 1.0, "storage": 0.0},
                      "parameter_server": {"gpu": 0.0, "cpu": 1.0, "memory": 1.0, "storage": 0.0},
                      "worker_price": {"bid": 1.0, "max": 2.0}, "parameter_server_price": {"bid": 1.0, "max": 2.0}},
    }

    async def kill_executor_of_c(nodes, logs):
        while not any(" round 1:" in l or l.startswith("round 1:") for l in logs):
            await asyncio.sleep(0.1)
        jobs = list(nodes["worker-c"].jobs.jobs.values())
        assert jobs, "worker-c runs no job"
        os.killpg(jobs[0].execution.pid, signal.SIGKILL)

    outcome, sink, logs, nodes = asyncio.run(run_loopback_job(
        tmp, repo="local/gpt2-tiny", write_model_config=lambda p: GPT2Config.tiny().to_json(p), job=job, num_workers=3,
        chaos=kill_executor_of_c, timeout_s=400,
        write_dataset=lambda d: write_synthetic_token_slices(d, num_slices=12, samples_per_slice=64, seq_len=32,
                                                             vocab_size=512)))
    assert outcome.finished, logs[-30:]
    assert outcome.rounds == 30 and outcome.removed_workers == ["worker-c"]
    assert outcome.joined_workers == ["worker-c"], (outcome.joined_workers, logs[-30:])
    assert any("joined a running job after round" in l for l in logs)
    assert sum("Finished training of" in l for l in logs) == 3


def test_default_style_lenet_job_with_image_preprocessor(tmp_path):
    """The reference's default job shape (scheduler_config.rs:52-105): LeNet image classification with an
    image preprocessor, parameter-server aggregation.  Slices carry raw `images`; the preprocessor turns
    them into `pixel_values`.  The synthetic classes are learnable, so the loss must fall."""
    from hypha_b200.executor.dataset import write_synthetic_image_slices
    from hypha_b200.models.lenet import LeNetConfig

    tmp = str(tmp_path)
    job = {
        "model": {"repository": "local/lenet", "filenames": ["config.json"], "type": "image-classification",
                  "input_names": ["pixel_values", "labels"]},
        "preprocessor": {"repository": "local/lenet", "filenames": ["preprocessor_config.json"], "type": "image",
                         "input_names": ["images"]},
        "dataset": {"dataset": "mnist-like"},
        "rounds": {"avg_samples_between_updates": 256, "update_rounds": 4, "max_batch_size": 32},
        "inner_optimizer": {"learning-rate": 3e-3},
        "outer_optimizer": {"learning-rate": 0.7, "momentum": 0.9},
        "resources": {"num_workers": 2, "worker": {"gpu": 0.125, "cpu": 1.0, "memory": 1.0, "storage": 0.0},
                      "parameter_server": {"gpu": 0.0, "cpu": 1.0, "memory": 1.0, "storage": 0.0},
                      "worker_price": {"bid": 1.0, "max": 2.0}, "parameter_server_price": {"bid": 1.0, "max": 2.0}},
    }

    def write_model(p):
        LeNetConfig().to_json(p)
        json.dump({"image_processor_type": "hypha-identity"}, open(os.path.join(os.path.dirname(p), "preprocessor_config.json"), "w"))

    outcome, sink, logs, nodes = asyncio.run(run_loopback_job(
        tmp, repo="local/lenet", write_model_config=write_model, job=job, timeout_s=300,
        write_dataset=lambda d: write_synthetic_image_slices(d, num_slices=8, samples_per_slice=128, pixel_key="images")))
    assert outcome.finished, logs[-20:]
    assert outcome.rounds == 4
    loss = {}
    for r in sink.records:
        if r["metric_name"] == "loss":
            loss.setdefault(r["worker_id"], {})[r["round"]] = r["value"]
    assert set(loss) == {"worker-a", "worker-b"}
    for w, v in loss.items():
        assert v[4] < v[1] * 0.8, (w, v)          # the model learns the synthetic classes


def test_job_checkpoint_and_resume(tmp_path):
    """Job-level checkpoint/resume (absent in the reference).  Run A trains 3 rounds and the
    aggregate worker saves momentum + the cumulative update U after every round.  Run B -- a new scheduler,
    new workers, same checkpoint directory, `resume = true` -- starts at round 3: the aggregate job loads
    the outer state, every worker is admitted like a newcomer (theta = theta_0(seed) + U) and three more
    rounds complete the six."""
    from safetensors.torch import load_file

    ckpt = str(tmp_path / "ckpt")

    def job(rounds, resume):
        return {
            "model": {"repository": "local/gpt2-tiny", "filenames": ["config.json"], "type": "causal-lm",
                      "input_names": ["input_ids", "labels"]},
            "preprocessor": None,
            "dataset": {"dataset": "toyset"},
            "rounds": {"avg_samples_between_updates": 40, "update_rounds": rounds, "max_batch_size": 4},
            "inner_optimizer": {"learning-rate": 3e-3},
            "outer_optimizer": {"learning-rate": 0.7, "momentum": 0.9},
            "checkpoint": {"dir": ckpt, "every": 1, "resume": resume},
            "resources": {"num_workers": 2, "worker": {"gpu": 1.0, "cpu": 1.0, "memory": 1.0, "storage": 0.0},
                          "parameter_server": {"gpu": 0.0, "cpu": 1.0, "memory": 1.0, "storage": 0.0},
                          "worker_price": {"bid": 1.0, "max": 2.0}, "parameter_server_price": {"bid": 1.0, "max": 2.0}},
        }

    def run(sub, j):
        d = str(tmp_path / sub)
        os.makedirs(d)
        return asyncio.run(run_loopback_job(
            d, repo="local/gpt2-tiny", write_model_config=lambda p: GPT2Config.tiny().to_json(p), job=j, timeout_s=240,
            write_dataset=lambda p: write_synthetic_token_slices(p, num_slices=6, samples_per_slice=16, seq_len=32,
                                                                 vocab_size=512)))

    a, _, logs_a, _ = run("a", job(3, False))
    assert a.finished and a.rounds == 3
    assert json.load(open(os.path.join(ckpt, "latest.json"))) == {"round": 3}
    assert sorted(os.listdir(ckpt)) == ["latest.json", "round-2", "round-3", "scheduler_state.json"]   # two newest kept
    st = json.load(open(os.path.join(ckpt, "scheduler_state.json")))                   # the scheduler's share
    assert st["round"] == 3 and len(st["slices"]) == 6 and any(p for _, p in st["slices"])
    assert {o for o, _ in st["slices"]} <= {"", "worker-a", "worker-b"}
    u3 = load_file(os.path.join(ckpt, "round-3", "cumulative_update"))
    assert u3 and all(torch.isfinite(v).all() for v in u3.values())

    b, sink_b, logs_b, _ = run("b", job(6, True))
    assert b.finished and b.rounds == 6, logs_b[-20:]
    assert any("resumed from round 3" in l for l in logs_b)
    assert sum("joined a running job after round 3" in l for l in logs_b) == 2
    assert sum("Finished training of 6 DiLoCo update rounds" in l for l in logs_b) == 2
    assert {r["round"] for r in sink_b.records} == {4, 5, 6}                          # only the new rounds ran
    assert json.load(open(os.path.join(ckpt, "latest.json"))) == {"round": 6}
    st6 = json.load(open(os.path.join(ckpt, "scheduler_state.json")))
    # the data scheduler went on from the restored state: same epoch bookkeeping, more slices used (or a new epoch)
    assert st6["round"] == 6 and (st6["epochs"], sum(p for _, p in st6["slices"])) != (st["epochs"], sum(p for _, p in st["slices"]))
    u6 = load_file(os.path.join(ckpt, "round-6", "cumulative_update"))
    assert any(not torch.equal(u6[k], u3[k]) for k in u3)                             # U kept accumulating


def test_performance_aware_rounds_give_the_fast_worker_more_batches(tmp_path):
    """rfc/2025-10-16_performance_aware_scheduling.md: H is not a constant.  worker-b is made 10x slower per
    step (`[[executors]].env`); the batch scheduler's sync-point projection, fed by the measured batch
    latencies, lets worker-a run several times more inner steps per round, and both still meet at every
    outer step."""
    import re

    tmp = str(tmp_path)
    job = {
        "model": {"repository": "local/gpt2-tiny", "filenames": ["config.json"], "type": "causal-lm",
                  "input_names": ["input_ids", "labels"]},
        "preprocessor": None,
        "dataset": {"dataset": "toyset"},
        "rounds": {"avg_samples_between_updates": 160, "update_rounds": 3, "max_batch_size": 4},
        "inner_optimizer": {"learning-rate": 3e-3},
        "outer_optimizer": {"learning-rate": 0.7, "momentum": 0.9},
        "aggregation": {"weighting": "samples"},        # unequal shares -> mean over samples, not over workers
        "resources": {"num_workers": 2, "worker": {"gpu": 1.0, "cpu": 1.0, "memory": 1.0, "storage": 0.0},
                      "parameter_server": {"gpu": 0.0, "cpu": 1.0, "memory": 1.0, "storage": 0.0},
                      "worker_price": {"bid": 1.0, "max": 2.0}, "parameter_server_price": {"bid": 1.0, "max": 2.0}},
    }
    outcome, sink, logs, nodes = asyncio.run(run_loopback_job(
        tmp, repo="local/gpt2-tiny", write_model_config=lambda p: GPT2Config.tiny().to_json(p), job=job, timeout_s=300,
        executor_env={"worker-a": {"HYPHA_EXECUTOR_STEP_DELAY_MS": "10"}, "worker-b": {"HYPHA_EXECUTOR_STEP_DELAY_MS": "100"}},
        write_dataset=lambda d: write_synthetic_token_slices(d, num_slices=8, samples_per_slice=32, seq_len=32,
                                                             vocab_size=512)))
    assert outcome.finished and outcome.rounds == 3, logs[-20:]
    steps = [int(m.group(1)) for l in logs for m in [re.search(r"round [23]: (\d+) inner steps", l)] if m]
    assert len(steps) == 4                                  # two workers x rounds 2 and 3 (round 1 warms the estimates up)
    fast, slow = max(steps), min(steps)
    assert fast >= 3 * slow, steps                          # the 10x faster worker runs several times more batches
    per_round = {}
    for l in logs:
        m = re.search(r"round (\d+): (\d+) inner steps", l)
        if m:
            per_round.setdefault(int(m.group(1)), []).append(int(m.group(2)))
    for r in (2, 3):                                        # together they cover the round's sample target
        assert 4 * sum(per_round[r]) >= 160 * 0.8, per_round
    # the aggregate worker weighted the pseudo-gradients by the samples behind them
    weights = [[float(x) for x in m.group(1).split(",")] for l in logs for m in [re.search(r"sample weights \[([^\]]+)\]", l)] if m]
    assert len(weights) == 3 and all(abs(sum(w) - 1.0) < 0.01 for w in weights), weights
    assert all(max(w) >= 0.7 for w in weights[1:]), weights      # rounds 2, 3: the fast worker dominates


def test_collective_job_survives_a_lost_rank(tmp_path):
    """Collective aggregation, three gloo ranks, worker-c dies after round 1.  `update` is a barrier at the
    scheduler over the CURRENT members; when the lost lease removes worker-c the barrier completes, its
    answer carries the new membership and a fresh rendezvous, the two survivors re-form their process group
    and finish the job."""
    tmp = str(tmp_path)
    job = {
        "model": {"repository": "local/gpt2-tiny", "filenames": ["config.json"], "type": "causal-lm",
                  "input_names": ["input_ids", "labels"]},
        "preprocessor": None,
        "dataset": {"dataset": "toyset"},
        "rounds": {"avg_samples_between_updates": 48, "update_rounds": 4, "max_batch_size": 4},
        "inner_optimizer": {"learning-rate": 3e-3},
        "outer_optimizer": {"learning-rate": 0.7, "momentum": 0.9},
        "aggregation": {"mode": "collective"},
        "resources": {"num_workers": 3, "worker": {"gpu": 1.0, "cpu": 1.0, "memory": 1.0, "storage": 0.0},
                      "parameter_server": {"gpu": 0.0, "cpu": 1.0, "memory": 1.0, "storage": 0.0},
                      "worker_price": {"bid": 1.0, "max": 2.0}, "parameter_server_price": {"bid": 1.0, "max": 2.0}},
    }

    async def kill_worker_c(nodes, logs):
        while not any("round 1:" in l for l in logs):
            await asyncio.sleep(0.1)
        nodes["worker-c"].stop.set()

    os.environ["HYPHA_EXECUTOR_DEVICE"] = "cpu"
    try:
        outcome, sink, logs, nodes = asyncio.run(run_loopback_job(
            tmp, repo="local/gpt2-tiny", write_model_config=lambda p: GPT2Config.tiny().to_json(p), job=job, num_workers=3,
            chaos=kill_worker_c, timeout_s=300,
            write_dataset=lambda d: write_synthetic_token_slices(d, num_slices=9, samples_per_slice=16, seq_len=32,
                                                                 vocab_size=512)))
    finally:
        os.environ.pop("HYPHA_EXECUTOR_DEVICE", None)
    assert outcome.finished, logs[-30:]
    assert outcome.rounds == 4 and outcome.removed_workers == ["worker-c"] and outcome.joined_workers == []
    assert sum("regrouped as rank" in l and "of 2" in l for l in logs) == 2, logs[-30:]
    assert sum("Finished training of 4 DiLoCo update rounds" in l for l in logs) == 2


def test_collective_job_replaces_a_lost_rank(tmp_path):
    """Collective aggregation with a spare worker: worker-c dies after round 1, the scheduler leases the
    spare, dispatches it with `collective.join`, and at the next `update` barrier everybody regroups: rank 0
    broadcasts the current model and outer momentum, the newcomer takes part in that outer step with a zero
    pseudo-gradient (excluded from the mean) and trains from then on."""
    tmp = str(tmp_path)
    job = {
        "model": {"repository": "local/gpt2-tiny", "filenames": ["config.json"], "type": "causal-lm",
                  "input_names": ["input_ids", "labels"]},
        "preprocessor": None,
        "dataset": {"dataset": "toyset"},
        "rounds": {"avg_samples_between_updates": 240, "update_rounds": 12, "max_batch_size": 4},
        "inner_optimizer": {"learning-rate": 3e-3},
        "outer_optimizer": {"learning-rate": 0.7, "momentum": 0.9},
        "aggregation": {"mode": "collective"},
        "resources": {"num_workers": 3, "worker": {"gpu": 1.0, "cpu": 1.0, "memory": 1.0, "storage": 0.0},
                      "parameter_server": {"gpu": 0.0, "cpu": 1.0, "memory": 1.0, "storage": 0.0},
                      "worker_price": {"bid": 1.0, "max": 2.0}, "parameter_server_price": {"bid": 1.0, "max": 2.0}},
    }

    async def kill_worker_c(nodes, logs):
        while not any("round 1:" in l for l in logs):
            await asyncio.sleep(0.1)
        nodes["worker-c"].stop.set()

    os.environ["HYPHA_EXECUTOR_DEVICE"] = "cpu"
    try:
        outcome, sink, logs, nodes = asyncio.run(run_loopback_job(
            tmp, repo="local/gpt2-tiny", write_model_config=lambda p: GPT2Config.tiny().to_json(p), job=job, num_workers=3,
            spare_workers=1, chaos=kill_worker_c, timeout_s=300,
            write_dataset=lambda d: write_synthetic_token_slices(d, num_slices=12, samples_per_slice=64, seq_len=32,
                                                                 vocab_size=512)))
    finally:
        os.environ.pop("HYPHA_EXECUTOR_DEVICE", None)
    assert outcome.finished, logs[-30:]
    assert outcome.rounds == 12 and outcome.removed_workers == ["worker-c"] and outcome.joined_workers == ["worker-d"]
    assert any("joined a running collective job at round" in l for l in logs), logs[-30:]
    assert sum("welcomed ['worker-d']" in l for l in logs) == 3                # both survivors and the newcomer
    rounds = {}
    for r in sink.records:
        if r["metric_name"] == "loss":
            rounds.setdefault(r["worker_id"], set()).add(r["round"])
    assert 12 in rounds["worker-a"] and 12 in rounds["worker-b"]
    assert len(rounds.get("worker-d", set())) >= 2 and 12 in rounds["worker-d"]     # it trained for real
    assert sum("Finished training of 12 DiLoCo update rounds" in l for l in logs) == 3


def test_transformers_auto_class_job(tmp_path):
    """A model family that is NOT one of ours (llama / gpt2 / lenet): `type = "sequence-classification"` with a
    BERT checkpoint resolves through the `transformers` auto-class table like the reference's executor
    (model.py:47-126), here without remote code.  The artifact carries config + weights, slices carry
    input_ids / attention_mask / labels, and the label is a function of the first token so the loss falls."""
    transformers = pytest.importorskip("transformers")
    from safetensors.torch import save_file

    tmp = str(tmp_path)
    job = {
        "model": {"repository": "local/bert-tiny", "filenames": ["config.json", "model.safetensors"],
                  "type": "sequence-classification", "input_names": ["input_ids", "attention_mask", "labels"]},
        "preprocessor": None,
        "dataset": {"dataset": "parity"},
        "rounds": {"avg_samples_between_updates": 128, "update_rounds": 3, "max_batch_size": 16},
        "inner_optimizer": {"learning-rate": 2e-3},
        "outer_optimizer": {"learning-rate": 0.7, "momentum": 0.9},
        "resources": {"num_workers": 2, "worker": {"gpu": 1.0, "cpu": 1.0, "memory": 1.0, "storage": 0.0},
                      "parameter_server": {"gpu": 0.0, "cpu": 1.0, "memory": 1.0, "storage": 0.0},
                      "worker_price": {"bid": 1.0, "max": 2.0}, "parameter_server_price": {"bid": 1.0, "max": 2.0}},
    }

    def write_model(p):
        torch.manual_seed(0)
        cfg = transformers.BertConfig(vocab_size=64, hidden_size=32, num_hidden_layers=2, num_attention_heads=2,
                                      intermediate_size=64, max_position_embeddings=32, num_labels=2,
                                      hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
        transformers.BertForSequenceClassification(cfg).save_pretrained(os.path.dirname(p))

    def write_dataset(d):
        os.makedirs(d, exist_ok=True)
        g = torch.Generator().manual_seed(1)
        for i in range(6):
            ids = torch.randint(0, 64, (64, 16), generator=g)
            save_file({"input_ids": ids, "attention_mask": torch.ones_like(ids), "labels": (ids[:, 0] % 2).contiguous()},
                      os.path.join(d, f"slice-{i:05d}.safetensors"))

    outcome, sink, logs, nodes = asyncio.run(run_loopback_job(
        tmp, repo="local/bert-tiny", write_model_config=write_model, job=job, write_dataset=write_dataset, timeout_s=300))
    assert outcome.finished, logs[-20:]
    assert outcome.rounds == 3
    loss = {}
    for r in sink.records:
        if r["metric_name"] == "loss":
            loss.setdefault(r["worker_id"], {})[r["round"]] = r["value"]
    assert set(loss) == {"worker-a", "worker-b"}
    for w, v in loss.items():
        assert v[3] < v[1], (w, v)


def test_collective_job_checkpoint_and_resume(tmp_path):
    """Checkpoint / resume in collective mode (CPU engine): after every outer step rank 0 writes the global
    model and the outer momentum; a second run with `resume = true` loads them on every rank, starts at the
    saved round and finishes the remaining rounds.  A checkpoint of the other aggregation mode is refused."""
    from safetensors.torch import load_file

    ckpt = str(tmp_path / "ckpt")

    def job(rounds, resume, mode="collective"):
        return {
            "model": {"repository": "local/gpt2-tiny", "filenames": ["config.json"], "type": "causal-lm",
                      "input_names": ["input_ids", "labels"]},
            "preprocessor": None,
            "dataset": {"dataset": "toyset"},
            "rounds": {"avg_samples_between_updates": 40, "update_rounds": rounds, "max_batch_size": 4},
            "inner_optimizer": {"learning-rate": 3e-3},
            "outer_optimizer": {"learning-rate": 0.7, "momentum": 0.9},
            "aggregation": {"mode": mode},
            "checkpoint": {"dir": ckpt, "every": 1, "resume": resume},
            "resources": {"num_workers": 2, "worker": {"gpu": 1.0, "cpu": 1.0, "memory": 1.0, "storage": 0.0},
                          "parameter_server": {"gpu": 0.0, "cpu": 1.0, "memory": 1.0, "storage": 0.0},
                          "worker_price": {"bid": 1.0, "max": 2.0}, "parameter_server_price": {"bid": 1.0, "max": 2.0}},
        }

    def run(sub, j):
        d = str(tmp_path / sub)
        os.makedirs(d)
        return asyncio.run(run_loopback_job(
            d, repo="local/gpt2-tiny", write_model_config=lambda p: GPT2Config.tiny().to_json(p), job=j, timeout_s=240,
            write_dataset=lambda p: write_synthetic_token_slices(p, num_slices=6, samples_per_slice=16, seq_len=32,
                                                                 vocab_size=512)))

    os.environ["HYPHA_EXECUTOR_DEVICE"] = "cpu"
    try:
        a, _, logs_a, _ = run("a", job(2, False))
        assert a.finished and a.rounds == 2, logs_a[-20:]
        assert json.load(open(os.path.join(ckpt, "latest.json"))) == {"round": 2}
        s2 = load_file(os.path.join(ckpt, "round-2", "collective_state"))
        assert any(k.startswith("prev/") for k in s2) and any(k.startswith("m/") for k in s2)
        assert sum("collective checkpoint of round" in l for l in logs_a) == 2            # one writer per round

        b, sink_b, logs_b, _ = run("b", job(4, True))
        assert b.finished and b.rounds == 4, logs_b[-20:]
        assert sum("resumed from the collective checkpoint of round 2" in l for l in logs_b) == 2   # every rank loads
        assert {r["round"] for r in sink_b.records} == {3, 4}
        assert json.load(open(os.path.join(ckpt, "latest.json"))) == {"round": 4}
        s4 = load_file(os.path.join(ckpt, "round-4", "collective_state"))
        moved = [k for k in s2 if k.startswith("prev/") and s2[k].dtype.is_floating_point and not torch.equal(s2[k], s4[k])]
        assert moved                                                                           # training went on

        with pytest.raises(Exception, match="collective mode"):
            run("c", job(6, True, mode="parameter-server"))
    finally:
        os.environ.pop("HYPHA_EXECUTOR_DEVICE", None)


def test_job_survives_two_lost_workers_one_after_the_other(tmp_path):
    """Cascaded failures in parameter-server mode: worker-c dies after round 1, worker-b after round 2; with
    `elastic.min_workers = 1` the last worker and the aggregate worker finish all rounds."""
    tmp = str(tmp_path)
    job = {
        "model": {"repository": "local/gpt2-tiny", "filenames": ["config.json"], "type": "causal-lm",
                  "input_names": ["input_ids", "labels"]},
        "preprocessor": None,
        "dataset": {"dataset": "toyset"},
        "rounds": {"avg_samples_between_updates": 48, "update_rounds": 4, "max_batch_size": 4},
        "inner_optimizer": {"learning-rate": 3e-3},
        "outer_optimizer": {"learning-rate": 0.7, "momentum": 0.9},
        "elastic": {"enabled": True, "min_workers": 1},
        "resources": {"num_workers": 3, "worker": {"gpu": 1.0, "cpu": 1.0, "memory": 1.0, "storage": 0.0},
                      "parameter_server": {"gpu": 0.0, "cpu": 1.0, "memory": 1.0, "storage": 0.0},
                      "worker_price": {"bid": 1.0, "max": 2.0}, "parameter_server_price": {"bid": 1.0, "max": 2.0}},
    }

    async def chaos(nodes, logs):
        for victim, after in (("worker-c", 1), ("worker-b", 2)):
            while not any(f"round {after}:" in l for l in logs):
                await asyncio.sleep(0.1)
            nodes[victim].stop.set()

    outcome, sink, logs, nodes = asyncio.run(run_loopback_job(
        tmp, repo="local/gpt2-tiny", write_model_config=lambda p: GPT2Config.tiny().to_json(p), job=job, num_workers=3,
        write_dataset=lambda d: write_synthetic_token_slices(d, num_slices=9, samples_per_slice=16, seq_len=32,
                                                             vocab_size=512), timeout_s=300, chaos=chaos))
    assert outcome.finished, logs[-25:]
    assert outcome.rounds == 4 and sorted(outcome.removed_workers) == ["worker-b", "worker-c"]
    assert sum("Finished training of 4 DiLoCo update rounds" in l for l in logs) == 1      # worker-a alone
    rounds_a = sorted(r["round"] for r in sink.records if r["worker_id"] == "worker-a" and r["metric_name"] == "loss")
    assert rounds_a == [1, 2, 3, 4]


def test_collective_job_survives_two_lost_ranks_down_to_one(tmp_path):
    """Collective aggregation under cascaded failures: three gloo ranks, worker-c dies after round 1, worker-b after
    round 2; the last rank goes on alone (its group has one member, the outer step needs no collective)."""
    tmp = str(tmp_path)
    job = {
        "model": {"repository": "local/gpt2-tiny", "filenames": ["config.json"], "type": "causal-lm",
                  "input_names": ["input_ids", "labels"]},
        "preprocessor": None,
        "dataset": {"dataset": "toyset"},
        "rounds": {"avg_samples_between_updates": 48, "update_rounds": 4, "max_batch_size": 4},
        "inner_optimizer": {"learning-rate": 3e-3},
        "outer_optimizer": {"learning-rate": 0.7, "momentum": 0.9},
        "aggregation": {"mode": "collective"},
        "elastic": {"enabled": True, "min_workers": 1},
        "resources": {"num_workers": 3, "worker": {"gpu": 1.0, "cpu": 1.0, "memory": 1.0, "storage": 0.0},
                      "parameter_server": {"gpu": 0.0, "cpu": 1.0, "memory": 1.0, "storage": 0.0},
                      "worker_price": {"bid": 1.0, "max": 2.0}, "parameter_server_price": {"bid": 1.0, "max": 2.0}},
    }

    async def chaos(nodes, logs):
        for victim, after in (("worker-c", 1), ("worker-b", 2)):
            while not any(f"round {after}:" in l for l in logs):
                await asyncio.sleep(0.1)
            nodes[victim].stop.set()

    os.environ["HYPHA_EXECUTOR_DEVICE"] = "cpu"
    try:
        outcome, sink, logs, nodes = asyncio.run(run_loopback_job(
            tmp, repo="local/gpt2-tiny", write_model_config=lambda p: GPT2Config.tiny().to_json(p), job=job, num_workers=3,
            chaos=chaos, timeout_s=300,
            write_dataset=lambda d: write_synthetic_token_slices(d, num_slices=9, samples_per_slice=16, seq_len=32,
                                                                 vocab_size=512)))
    finally:
        os.environ.pop("HYPHA_EXECUTOR_DEVICE", None)
    assert outcome.finished, logs[-30:]
    assert outcome.rounds == 4 and sorted(outcome.removed_workers) == ["worker-b", "worker-c"]
    assert sum("Finished training of 4 DiLoCo update rounds" in l for l in logs) == 1
    assert any("regrouped as rank 0 of 1" in l for l in logs), [l for l in logs if "regroup" in l]
