"""Bring your own executor: the contract between a worker and the program it starts for a train job, in its smallest
form (the reference documents the same contract in executors/accelerate/README.md; `hypha_b200.executor.training` is
the full implementation).

A worker runs whatever its configuration names for the executor class of the job:

    [[executors]]
    class = "train"
    name = "diloco-transformer"
    runtime = "process"
    cmd = "python"
    args = ["examples/custom_executor.py", "--socket", "{SOCKET_PATH}", "--work-dir", "{WORK_DIR}", "--job", "{JOB_JSON}"]

`{SOCKET_PATH}` is the job bridge (HTTP over a Unix socket, `Session` below), `{WORK_DIR}` the job's private directory
(every path sent to the bridge is relative to it), `{JOB_JSON}` the job specification.  The loop below is the DiLoCo
worker protocol without a model: fetch the artifact, report one `status` per batch and obey the scheduler's answers
(`Continue`, `ScheduleUpdate{counter}` = "synchronise after `counter` more batches", `Done`), then send the
pseudo-gradient file to the aggregating peers, wait for the update, report `update-received`, report metrics.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main() -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--socket", required=True)
    ap.add_argument("--work-dir", required=True)
    ap.add_argument("--job", required=True, help="the job specification (JSON text)")
    args = ap.parse_args()

    import torch
    from safetensors.torch import save_file

    from hypha_b200.executor.api import Session

    job = json.loads(args.job)
    cfg = job["executor"]["config"]
    with Session(args.socket) as bridge:
        files = bridge.fetch(cfg["model"]["artifact"])          # -> [{"path": "artifacts/...", "size": n}, ...]
        print(f"job {job['job_id']}: fetched {[f['path'] for f in files]}", flush=True)
        theta = torch.zeros(16)                                 # stands in for the model
        round_no = 0
        while True:
            # ---- inner loop: one status per batch; the scheduler decides when the round ends -----------------------
            countdown = None
            while countdown is None or countdown > 0:
                theta += 0.01 * torch.randn_like(theta)         # "one optimizer step"
                answer = bridge.send_status({"status": {"batch_size": cfg["batch_size"]}})
                if answer["type"] == "Done":
                    return 0
                if answer["type"] == "ScheduleUpdate":
                    countdown = int(answer["counter"])
                elif countdown is not None:
                    countdown -= 1
            # ---- outer step: pseudo-gradient out, update in ------------------------------------------------------
            save_file({"theta": theta.clone()}, os.path.join(args.work_dir, "delta.safetensors"))
            with bridge.receive(cfg["results"], "incoming") as updates:       # subscribe BEFORE sending
                bridge.send_resource(cfg["updates"], "delta.safetensors")
                for event in updates:                                         # {"path", "size", "from_peer"}
                    print(f"round {round_no}: update {event['path']} from {event['from_peer']}", flush=True)
                    break
            if bridge.send_status("update-received")["type"] == "Done":
                return 0
            bridge.send_status({"metrics": {"round": round_no, "metrics": {"loss": float(theta.norm())}}})
            round_no += 1


if __name__ == "__main__":
    raise SystemExit(main())
