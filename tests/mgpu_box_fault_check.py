"""The product path under a REAL process death (BASELINE config 4; the reference's scheduler
exits on the first failed worker, crates/scheduler/src/bin/hypha-scheduler.rs:400-405).

    python tests/mgpu_box_fault_check.py [--workers 3] [--model llama-small]

gateway + data node + `workers` train workers (one per GPU) + scheduler in this process over the loopback
fabric; a `collective` DiLoCo job whose executors (one OS process per GPU) run the fused peer-memory outer
step.  After round 1 the executor PROCESS of worker-b is killed with SIGKILL while the others keep training:

  * its worker reports JobStatus Failed / the scheduler stops renewing the lease -> the rank is dropped,
    the survivors' `update` barrier returns the shrunk membership, `DiLoCoOuter.drop_ranks` adopts the shard
    (with the owner's fp32 theta/momentum reconstructed, see parallel/diloco.py) and the outer step runs on
    the survivors;
  * `elastic.replace` leases a worker again (the same worker is the cheapest offer: its GPU is free now),
    dispatches a `collective.join` job with the dead rank's slot; the NEW process maps the survivors' buffers
    over CUDA IPC (`PeerGroup.joining` / `rebind`), is admitted at the next round boundary
    (`DiLoCoOuter.admit`) and pulls the weights over NVLink;
  * the job finishes all rounds with the full membership.

Prints one JSON line; exit status 0 iff every check holds."""
import argparse
import asyncio
import json
import os
import signal
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--workers", type=int, default=3)
    ap.add_argument("--model", default="llama-small")
    ap.add_argument("--rounds", type=int, default=8)
    ap.add_argument("--seq-len", type=int, default=512)
    ap.add_argument("--batch-size", type=int, default=2)
    ap.add_argument("--steps-per-round", type=int, default=8)
    ap.add_argument("--nvls", default="0", help="HYPHA_NVLS for the executors (a joining process needs re-bindable buffers)")
    a = ap.parse_args()

    from hypha_b200.executor.dataset import write_synthetic_token_slices
    from hypha_b200.local_cluster import run_loopback_job
    from hypha_b200.models.llama import get_config

    os.environ["HYPHA_EXECUTOR_DEVICE"] = "cuda"
    os.environ["HYPHA_GPU_JOIN"] = "1"
    os.environ["HYPHA_NVLS"] = a.nvls
    cfg = get_config(a.model, max_seq_len=max(a.seq_len, 8))
    repo = f"local/{a.model}"
    n = a.workers
    job = {
        "model": {"repository": repo, "filenames": ["config.json"], "type": "causal-lm", "input_names": ["input_ids", "labels"]},
        "preprocessor": None, "dataset": {"dataset": "synthetic-tokens"},
        "rounds": {"avg_samples_between_updates": n * a.batch_size * a.steps_per_round, "update_rounds": a.rounds,
                   "max_batch_size": a.batch_size},
        "inner_optimizer": {"learning-rate": 1e-4}, "outer_optimizer": {"learning-rate": 0.7, "momentum": 0.9},
        "aggregation": {"mode": "collective"},
        "resources": {"num_workers": n, "worker": {"gpu": 1.0, "cpu": 1.0, "memory": 1.0, "storage": 0.0},
                      "parameter_server": {"gpu": 0.0, "cpu": 1.0, "memory": 1.0, "storage": 0.0},
                      "worker_price": {"bid": 1.0, "max": 2.0}, "parameter_server_price": {"bid": 1.0, "max": 2.0}},
    }
    t_kill = {}

    async def kill_executor_of_b(nodes, logs):
        while not any(" round 1:" in l for l in logs):
            await asyncio.sleep(0.05)
        jobs = list(nodes["worker-b"].jobs.jobs.values())
        assert jobs, "worker-b runs no job"
        os.killpg(jobs[0].execution.pid, signal.SIGKILL)
        t_kill["t"] = time.time()
        t_kill["pid"] = jobs[0].execution.pid

    tmp = tempfile.mkdtemp(prefix="hypha-box-fault-")
    t0 = time.time()
    outcome, sink, logs, nodes = asyncio.run(run_loopback_job(
        tmp, repo=repo, write_model_config=lambda p: cfg.to_hf_json(p, seed=5), job=job, num_workers=n, timeout_s=900,
        chaos=kill_executor_of_b,
        write_dataset=lambda d: write_synthetic_token_slices(d, num_slices=max(2 * n, 4), samples_per_slice=64,
                                                             seq_len=a.seq_len, vocab_size=cfg.vocab_size)))
    rounds = {}
    for r in sink.records:
        if r["metric_name"] == "loss":
            rounds.setdefault(r["worker_id"], []).append(int(r["round"]))
    checks = {
        "finished": bool(outcome.finished),
        "all_rounds": outcome.rounds == a.rounds,
        "rank_dropped": outcome.removed_workers == ["worker-b"],
        "replacement_joined": list(outcome.joined_workers) == ["worker-b"],
        "survivors_continued_without_it": any("continuing without ranks" in l or "continued without ranks" in l for l in logs),
        "admitted_into_fused_engine": any("admitted into the fused engine" in l for l in logs),
        "newcomer_trained": any("joined a running collective job at round" in l for l in logs)
                            and max(rounds.get("worker-b", [0])) == a.rounds,
        "everyone_finished": sum(f"Finished training of {a.rounds} DiLoCo update rounds" in l for l in logs) == n,
    }
    out = {"ok": all(checks.values()), "checks": checks, "workers": n, "model": a.model, "wall_s": round(time.time() - t0, 1),
           "killed_pid": t_kill.get("pid"), "rounds_reported": {k: sorted(set(v)) for k, v in rounds.items()}}
    print(json.dumps(out))
    if not out["ok"]:
        print("\n".join(logs[-60:]), file=sys.stderr)
    return 0 if out["ok"] else 1


if __name__ == "__main__":
    sys.exit(main())
