#!/usr/bin/env python
"""BASELINE config 1: GPT-2 small (124 M) DiLoCo, 2 workers on CPU, through the loopback
gateway / scheduler / workers / data node (no GPU).  H ~ 10 inner steps per round by construction
(avg_samples_between_updates = 10 * batch * workers).  Prints one JSON line."""
import argparse
import asyncio
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypha_b200.executor.dataset import write_synthetic_token_slices  # noqa: E402
from hypha_b200.local_cluster import run_loopback_job  # noqa: E402
from hypha_b200.models.gpt2 import GPT2Config  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=2)
    ap.add_argument("--seq-len", type=int, default=64)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--inner-steps", type=int, default=10)
    ap.add_argument("--tiny", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    cfg = GPT2Config.tiny() if a.tiny else GPT2Config.small()
    job = {
        "model": {"repository": "local/gpt2", "filenames": ["config.json"], "type": "causal-lm",
                  "input_names": ["input_ids", "labels"]},
        "preprocessor": None, "dataset": {"dataset": "synthetic-tokens"},
        "rounds": {"avg_samples_between_updates": a.inner_steps * a.batch * 2, "update_rounds": a.rounds,
                   "max_batch_size": a.batch},
        "inner_optimizer": {"learning-rate": 3e-4}, "outer_optimizer": {"learning-rate": 0.7, "momentum": 0.9},
        "resources": {"num_workers": 2, "worker": {"gpu": 1.0, "cpu": 1.0, "memory": 1.0, "storage": 0.0},
                      "parameter_server": {"gpu": 0.0, "cpu": 1.0, "memory": 1.0, "storage": 0.0},
                      "worker_price": {"bid": 1.0, "max": 2.0}, "parameter_server_price": {"bid": 1.0, "max": 2.0}},
    }
    with tempfile.TemporaryDirectory(prefix="hb-plumb-", dir="/tmp") as tmp:
        t0 = time.time()
        outcome, sink, logs, _ = asyncio.run(run_loopback_job(
            tmp, repo="local/gpt2", write_model_config=lambda p: cfg.to_json(p), job=job,
            write_dataset=lambda d: write_synthetic_token_slices(d, 8, 32, a.seq_len, cfg.vocab_size),
            timeout_s=3600, omp_threads=max(1, (os.cpu_count() or 4) // 2),
            log=(lambda m: print(m, flush=True)) if a.verbose else None))
        wall = time.time() - t0
    losses = {}
    for r in sink.records:
        losses.setdefault(r["worker_id"], {})[r["round"]] = round(r["value"], 4)
    print(json.dumps({"config": "GPT-2 %s DiLoCo, 2 CPU workers, loopback control plane" % ("tiny" if a.tiny else "small"),
                      "params": cfg.num_params(), "finished": outcome.finished, "rounds": outcome.rounds,
                      "inner_steps_per_round": a.inner_steps, "wall_s": round(wall, 1), "loss_by_worker": losses,
                      "workers": outcome.workers, "parameter_server": outcome.parameter_server}))
    return 0 if outcome.finished else 1


if __name__ == "__main__":
    sys.exit(main())
