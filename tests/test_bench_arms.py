"""bench.py arms that can be exercised without a GPU.

The reference arm (``--impl reference``) is the stock re-creation of the reference's training loop
around the reference's own utils/dataset code (bench/reference_arm.py); on CPU it runs only as a
plumbing check of the JSON contract (``HYPHA_BENCH_ALLOW_CPU=1``, tiny model, wall-clock timing)."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra, env_extra=None, timeout=600):
    env = dict(os.environ, HYPHA_BENCH_ALLOW_CPU="1", CUDA_VISIBLE_DEVICES="")
    env.update(env_extra or {})
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py")] + extra, env=env, capture_output=True,
                         text=True, timeout=timeout)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout
    return json.loads(lines[0])


def test_reference_arm_json_contract_cpu():
    r = _run(["--impl", "reference", "--model", "llama-tiny", "--seq-len", "32", "--micro-batch", "2",
              "--steps", "3", "--warmup", "3", "--inner-steps", "2"])
    assert r["impl"] == "reference" and r["reference_class"] == "stock-recreation"
    assert r["steps"] == 3 and r["warmup"] == 3 and r["n_gpus"] == 1
    assert r["value"] > 0 and r["e2e"]["value"] > 0
    assert r["e2e"]["h2d_bytes_per_step"] == 2 * 2 * 32 * 8          # input_ids + labels, int64
    assert r["config"]["outer_steps_in_timed_region"] >= 1
    assert "hypha_b200" not in r["config"]["model"]


def test_reference_arm_two_workers_gloo():
    env = dict(os.environ, HYPHA_BENCH_ALLOW_CPU="1", CUDA_VISIBLE_DEVICES="")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(REPO, "bench.py"), "--impl", "stock", "--gpus", "2",
           "--model", "llama-tiny", "--seq-len", "32", "--micro-batch", "1", "--steps", "3", "--warmup", "3",
           "--inner-steps", "2"]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    r = json.loads(lines[0])
    assert r["n_gpus"] == 2 and r["config"]["global_batch"] == 2 and r["outer_step_ms"] is not None


def test_product_arm_refuses_cpu():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--steps", "1"], env=env,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 2 and "sm_100a" in out.stderr
