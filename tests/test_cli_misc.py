import asyncio
import json
import os
import tomllib

import pytest

from hypha_b200 import certutil as cu
from hypha_b200 import cli, config as C, roles
from hypha_b200.cli import certutil as cu_cli
from hypha_b200.worker import InvalidPath, nesterov_step, safe_join, substitute_placeholders


def test_init_writes_documented_toml_and_job_injection(tmp_path):
    out = tmp_path / "sched.toml"
    job = {"rounds": {"update_rounds": 7}, "resources": {"num_workers": 8}}
    assert cli.main("scheduler", ["init", "-o", str(out), "--job", json.dumps(job), "-n", "sched1"]) == 0
    d = tomllib.loads(out.read_text())
    assert d["scheduler"]["job"]["rounds"]["update_rounds"] == 7 and d["scheduler"]["job"]["resources"]["num_workers"] == 8
    assert d["cert_pem"] == "sched1-cert.pem" and "# " in out.read_text()
    for role in ("gateway", "worker", "data"):
        p = tmp_path / f"{role}.toml"
        assert cli.main(role, ["init", "-o", str(p)]) == 0 and tomllib.loads(p.read_text())
        assert "## hypha-%s run" % role in cli.cli_markdown(role)


def test_run_rejects_bad_config(tmp_path, capsys):
    p = tmp_path / "w.toml"
    p.write_text('gateway_addresses = ["tcp://10.0.0.1:1"]\n')
    assert cli.main("worker", ["run", "-c", str(p)]) == 2
    assert "excluded range" in capsys.readouterr().err


def test_probe_healthy_and_unreachable():
    async def main():
        gw_cfg = C.builder(C.GatewayConfig).with_overrides({"listen_addresses": ["tcp://127.0.0.1:0"]}).build()
        fut = asyncio.get_event_loop().create_future()
        t = asyncio.ensure_future(roles.run_gateway(gw_cfg, started=fut))
        node = await fut
        addr = node.network.listen_addrs[0]
        assert await roles.probe(addr)
        node.stop.set()
        await t
        assert not await roles.probe(addr, timeout=0.5)
    asyncio.run(main())


def test_certutil_chain(tmp_path):
    d = str(tmp_path)
    assert cu_cli.main(["root", "-o", "acme", "-d", d]) == 0
    assert cu_cli.main(["org", "--root-cert", f"{d}/acme-root-cert.pem", "--root-key", f"{d}/acme-root-key.pem",
                        "-o", "acme", "-d", d]) == 0
    assert cu_cli.main(["node", "--ca-cert", f"{d}/acme-ca-cert.pem", "--ca-key", f"{d}/acme-ca-key.pem",
                        "-n", "worker-1", "-d", d]) == 0
    assert cu.verify_chain(f"{d}/worker-1-cert.pem")
    assert oct(os.stat(f"{d}/worker-1-key.pem").st_mode & 0o777) == "0o600"
    assert b"PRIVATE KEY" in open(f"{d}/worker-1-key.pem", "rb").read()
    cfg = C.builder(C.WorkerConfig).with_overrides({"cert_pem": f"{d}/worker-1-cert.pem", "key_pem": f"{d}/worker-1-key.pem",
                                                    "trust_pem": f"{d}/worker-1-trust.pem"}).build()
    assert C.load_tls(cfg)["trust_pem"].startswith(b"-----BEGIN CERTIFICATE")


def test_substitutes_all_known_tokens():   # executor/process.rs:207-241
    out = substitute_placeholders(["--socket", "{SOCKET_PATH}", "--work-dir={WORK_DIR}", "{JOB_JSON}", "{UNKNOWN}"],
                                  "/s.sock", "/w", '{"a":1}')
    assert out == ["--socket", "/s.sock", "--work-dir=/w", '{"a":1}', "{UNKNOWN}"]


def test_safe_join_rules():                # bridge.rs:330-346
    assert safe_join("/work", "artifacts/x.bin") == "/work/artifacts/x.bin"
    for bad in ("/etc/passwd", "../x", "a/../../x", ""):
        with pytest.raises(InvalidPath):
            safe_join("/work", bad)


def test_nesterov_reference_vector():      # parameter_server.rs:452-524
    import torch

    m, u = nesterov_step({"w": torch.full((5,), 0.5)}, None, 0.1, 0.7)
    assert torch.allclose(u["w"], torch.full((5,), 0.085))
    m, u = nesterov_step({"w": torch.tensor([0.1, 0.2, 0.3, 0.4, 0.5])}, m, 0.1, 0.7)
    assert torch.allclose(u["w"], torch.tensor([0.0415, 0.0585, 0.0755, 0.0925, 0.1095]), atol=1e-7)


def test_aim_driver_jsonl_fallback(tmp_path, monkeypatch):
    from fastapi.testclient import TestClient

    import hypha_b200.drivers.aim_driver as drv

    monkeypatch.setattr(drv, "JSONL", str(tmp_path / "m.jsonl"))
    c = TestClient(drv.app)
    r = c.post("/status", json={"worker_id": "w1", "round": 3, "metric_name": "loss", "value": 1.5})
    assert r.status_code == 200
    assert json.loads((tmp_path / "m.jsonl").read_text())["value"] == 1.5
    assert c.post("/status", json={"worker_id": "w1"}).status_code == 422


def test_hypha_box_runs_a_collective_job_on_cpu(capsys):
    """`hypha-box run`: gateway + data + 2 workers + scheduler in one process, collective aggregation
    over gloo, one JSON line with per-worker per-round metrics."""
    import json

    from hypha_b200.cli import box

    rc = box.main(["run", "--model", "gpt2-tiny", "--workers", "2", "--device", "cpu", "--rounds", "2",
                   "--samples-per-round", "32", "--batch-size", "4", "--seq-len", "32", "--timeout", "240"])
    out = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert rc == 0 and out["finished"] and out["rounds"] == 2 and out["aggregation"] == "collective"
    assert sorted(out["metrics"]) == ["worker-a", "worker-b"]
    assert set(out["metrics"]["worker-a"]["2"]) == {"loss", "samples_or_tokens_per_s", "inner_step_ms", "outer_step_ms",
                                                     "inner_steps", "tokens"}


def test_worker_init_detects_resources_and_pins_a_device(tmp_path):
    from hypha_b200 import cli
    from hypha_b200 import config as C

    out = tmp_path / "w.toml"
    assert cli.main("worker", ["init", "--detect", "--device", "2", "-o", str(out)]) == 0
    cfg = C.builder(C.WorkerConfig).with_toml(str(out)).build()
    assert cfg.device == 2 and cfg.resources.cpu >= 1 and cfg.resources.memory >= 1
    assert cfg.resources.gpu in (0.0, 1.0)            # pinned to one device (0 on a CPU-only host)


def test_replicated_nesterov_matches_the_reference_vector():
    """Single-rank collective engine (no process group): the outer step is the reference's Nesterov
    (parameter_server.rs:452-524): lr=.1 mu=.7, g0=.5 -> u0=.085; g1=[.1..5] -> u1=[.0415 .. .1095]."""
    import torch

    from hypha_b200.executor.training import ReplicatedNesterov

    class _Loop:
        flat = False

        def __init__(self):
            self.model = torch.nn.Linear(5, 1, bias=False)
            with torch.no_grad():
                self.model.weight.zero_()
            self.prev = {k: v.detach().clone() for k, v in self.model.state_dict().items()}

    loop, outer = _Loop(), ReplicatedNesterov(0.1, 0.7)
    with torch.no_grad():
        loop.model.weight.add_(0.5)                       # pseudo-gradient g0 = 0.5
    outer.step(loop, 1)
    assert torch.allclose(loop.model.weight, torch.full((1, 5), 0.085), atol=1e-7)
    with torch.no_grad():
        loop.model.weight.add_(torch.tensor([[0.1, 0.2, 0.3, 0.4, 0.5]]))
    outer.step(loop, 1)
    want = 0.085 + torch.tensor([[0.0415, 0.0585, 0.0755, 0.0925, 0.1095]])
    assert torch.allclose(loop.model.weight, want, atol=1e-6)
    assert torch.equal(loop.prev["weight"], loop.model.weight)


def test_hypha_box_single_worker_on_cpu(capsys, tmp_path):
    """... and the way out: `--checkpoint-dir` + `--export` leave a model directory that `transformers` loads, holding
    weights that moved away from the seeded initialisation; `hypha-box export` does the same from the checkpoint."""
    import json

    import torch
    import transformers

    from hypha_b200.cli import box
    from hypha_b200.models.gpt2 import GPT2Config

    ck, out_dir = str(tmp_path / "ckpt"), str(tmp_path / "trained")
    rc = box.main(["run", "--model", "gpt2-tiny", "--workers", "1", "--device", "cpu", "--rounds", "2",
                   "--samples-per-round", "16", "--batch-size", "4", "--seq-len", "32", "--timeout", "240",
                   "--lr", "1e-2", "--checkpoint-dir", ck, "--export", out_dir])
    out = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert rc == 0 and out["finished"] and out["rounds"] == 2 and list(out["metrics"]) == ["worker-a"]
    assert out["exported"]["round"] == 2 and out["exported"]["kind"] == "collective"
    hf = transformers.GPT2LMHeadModel.from_pretrained(out_dir).eval()
    with torch.no_grad():
        assert torch.isfinite(hf(torch.randint(0, 100, (1, 16))).logits).all()
    # the same through the stand-alone command, from the checkpoint and the model config
    mdir = tmp_path / "model"
    mdir.mkdir()
    GPT2Config.tiny().to_json(str(mdir / "config.json"))
    assert box.main(["export", "--checkpoint", ck, "--model", str(mdir), "-o", str(tmp_path / "again"), "--dtype", "fp32"]) == 0
    info = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert info["round"] == 2 and info["parameters"] == sum(p.numel() for p in hf.parameters())
    again = transformers.GPT2LMHeadModel.from_pretrained(str(tmp_path / "again"))
    from hypha_b200.executor.model import get_model

    init = get_model(str(mdir), "causal-lm", "cpu").state_dict()
    moved = 0
    for k, v in hf.state_dict().items():
        assert torch.equal(v, again.state_dict()[k]), k
        moved += int(k in init and not torch.allclose(v, init[k].float(), atol=1e-6))
    assert moved > 10                                   # training happened to what was exported
    assert box.main(["export", "--checkpoint", str(tmp_path / "nowhere"), "--model", str(mdir), "-o", str(tmp_path / "x")]) == 2
    assert box.main(["run", "--model", "gpt2-tiny", "--device", "cpu", "--export", str(tmp_path / "y")]) == 2


def test_status_pipe_runs_exactly_the_projected_number_of_batches(monkeypatch):
    """Asynchronous status reports (quirk Q10): the answer to batch k-1 is read after batch k, and a
    `ScheduleUpdate{counter}` is corrected for that batch, so async and sync modes stop at the same batch."""
    from hypha_b200.executor.training import StatusPipe

    class FakeSession:
        def __init__(self, schedule_at, counter):
            self.n, self.schedule_at, self.counter, self.sent = 0, schedule_at, counter, []

        def send_status(self, msg):
            self.n += 1
            self.sent.append(msg)
            if self.n == self.schedule_at:
                return {"type": "ScheduleUpdate", "counter": self.counter}
            return {"type": "Continue", "counter": 0}

    def run(sync, schedule_at, c):
        monkeypatch.setenv("HYPHA_SYNC_STATUS", "1" if sync else "0")
        sess = FakeSession(schedule_at, c)
        pipe, counter, batches = StatusPipe(sess), -1, 0
        while counter != 0:
            batches += 1
            counter = pipe.report(4, counter)
            assert batches < 100
        pipe.drain()
        return batches, sess.n

    for schedule_at, c in ((1, 3), (2, 1), (3, 5), (4, 2)):
        sync_batches, sync_reports = run(True, schedule_at, c)
        async_batches, async_reports = run(False, schedule_at, c)
        assert sync_batches == schedule_at + c
        assert async_batches == sync_batches and async_reports == async_batches      # one report per batch
    # counter 0 answered one batch late: a single batch of overshoot, never a hang
    assert run(True, 2, 0)[0] == 2 and run(False, 2, 0)[0] == 3


def test_lr_schedules_match_transformers():
    """constant / cosine-with-warmup / linear-with-warmup / wsd (utils.py:90-102) against the Hugging Face
    schedulers the reference executor calls; `decay_step` (the reference's typo, quirk Q4) is an alias."""
    import torch
    import transformers

    from hypha_b200.parallel.diloco import make_lr_lambda, shard_bounds

    def hf(factory, n, **kw):
        opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1.0)
        sched = factory(opt, **kw)
        out = []
        for _ in range(n):
            out.append(opt.param_groups[0]["lr"])
            opt.step()
            sched.step()
        return out

    n = 60
    ours = [make_lr_lambda({"type": "cosine-with-warmup", "warmup_steps": 5, "training_steps": 50})(s) for s in range(n)]
    ref = hf(transformers.get_cosine_schedule_with_warmup, n, num_warmup_steps=5, num_training_steps=50)
    assert all(abs(a - b) < 1e-6 for a, b in zip(ours[:51], ref[:51]))
    assert all(v == 0.0 for v in ours[50:])      # past the end we stay at 0 (HF's cosine starts rising again)
    ours = [make_lr_lambda({"type": "linear-with-warmup", "warmup_steps": 5, "training_steps": 50})(s) for s in range(n)]
    ref = hf(transformers.get_linear_schedule_with_warmup, n, num_warmup_steps=5, num_training_steps=50)
    assert all(abs(a - b) < 1e-6 for a, b in zip(ours, ref))
    wsd = {"type": "wsd", "warmup_steps": 5, "stable_steps": 20, "decay_steps": 10}
    ours = [make_lr_lambda(wsd)(s) for s in range(n)]
    ref = hf(transformers.get_wsd_schedule, n, num_warmup_steps=5, num_stable_steps=20, num_decay_steps=10)
    assert all(abs(a - b) < 1e-6 for a, b in zip(ours, ref)), list(zip(ours, ref))[:40]
    typo = dict(wsd)
    typo["decay_step"] = typo.pop("decay_steps")
    assert [make_lr_lambda(typo)(s) for s in range(n)] == ours
    assert all(make_lr_lambda(None)(s) == 1.0 for s in range(5))
    # ZeRO-1 shard layout of the flat vector: aligned, disjoint, covering
    for total, world in ((8 * 1024 + 8, 8), (1 << 20, 3), (4096, 8)):
        spans = [shard_bounds(total, world, r) for r in range(world)]
        assert spans[0][0] == 0 and sum(l for _, l in spans) == total
        assert all(b % 1024 == 0 for b, _ in spans if b < total)
        assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] or spans[i + 1][1] == 0 for i in range(world - 1))


def test_cli_reference_is_up_to_date():
    """docs/cli.md is generated from the parsers (`tools/gen_cli_docs.py`); a flag added without regenerating
    the reference fails here."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "gen_cli_docs.py"), "--check"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr


def test_certutil_creates_the_output_directory_and_reports_missing_files(tmp_path, capsys):
    from hypha_b200.cli import certutil as cli_cu

    d = str(tmp_path / "pki" / "nested")                      # does not exist yet
    assert cli_cu.main(["root", "-o", "Acme", "-d", d]) == 0
    assert sorted(os.listdir(d)) == ["Acme-root-cert.pem", "Acme-root-key.pem"]
    assert cli_cu.main(["node", "--ca-cert", os.path.join(d, "missing.pem"), "--ca-key", os.path.join(d, "Acme-root-key.pem"),
                        "-n", "w0", "-d", d]) == 1
    assert "hypha-certutil: error:" in capsys.readouterr().err


def test_bench_contract_on_a_box_without_gpus():
    """`bench.py --impl reference` must print ONE JSON line saying why the reference arm is unavailable and exit
    0 (the driver runs it first on every box); the default arm needs a GPU and must say so instead of falling
    back to anything slower."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "2",
                        "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    out = json.loads(lines[0])
    assert out["impl"] == "reference" and isinstance(out["unavailable"], str) and "\n" not in out["unavailable"]
    import torch

    if not torch.cuda.is_available():
        r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1", "--warmup", "1"],
                           capture_output=True, text=True, timeout=300, cwd=root)
        assert r.returncode != 0 and "CUDA" in (r.stderr + r.stdout)


def test_replicated_nesterov_sample_weighted_mean_over_two_gloo_ranks():
    """`aggregation.weighting = "samples"` in the collective engine: two ranks (threads, owned gloo groups) with
    pseudo-gradients 1.0 and 4.0 after 10 and 30 samples -> mean over samples 3.25 (mean over ranks: 2.5); a rank
    that processed nothing (a newcomer) gets weight zero."""
    import socket
    import threading

    import torch

    from hypha_b200.executor.training import ReplicatedNesterov, _Group

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()

    class _Loop:
        flat, device = False, "cpu"

        def __init__(self, delta):
            self.model = torch.nn.Linear(4, 1, bias=False)
            with torch.no_grad():
                self.model.weight.zero_()
            self.prev = {k: v.detach().clone() for k, v in self.model.state_dict().items()}
            with torch.no_grad():
                self.model.weight.add_(delta)

    out = {}

    def rank(r, delta, samples_by_round):
        group = _Group("gloo", f"tcp://127.0.0.1:{port}", r, 2)
        loop, outer = _Loop(delta), ReplicatedNesterov(1.0, 0.0)           # lr 1, mu 0: theta moves by the mean
        res = []
        for rnd, n in enumerate(samples_by_round):
            outer.step(loop, group, rnd, 0, n)
            res.append(float(loop.model.weight.detach()[0, 0]))
            with torch.no_grad():
                loop.model.weight.add_(delta)                             # the next round's local progress
        out[r] = res

    ts = [threading.Thread(target=rank, args=(0, 1.0, [10, 0, None])), threading.Thread(target=rank, args=(1, 4.0, [30, 8, None]))]
    for t in ts:
        t.start()
    for t in ts:
        t.join(120)
    assert not any(t.is_alive() for t in ts)
    for r in (0, 1):
        assert out[r][0] == pytest.approx(0.25 * 1.0 + 0.75 * 4.0)          # round 0: mean over samples
        assert out[r][1] == pytest.approx(out[r][0] + 4.0)                  # round 1: rank 0 had no samples -> weight 0
        assert out[r][2] == pytest.approx(out[r][1] + 2.5)                  # round 2: no counts -> mean over ranks
    assert out[0] == out[1]


def test_process_executor_escalates_from_sigterm_to_sigkill(tmp_path, monkeypatch):
    """Cancellation of a process executor (process.rs:148-187): SIGTERM to the process group, a grace period,
    then SIGKILL; a job that honours SIGTERM ends within the grace period; the work dir is removed either way."""
    import signal
    import sys
    import time

    from hypha_b200 import messages as M
    from hypha_b200.transport import Network
    from hypha_b200.worker import Connector
    from hypha_b200.worker import jobs as J

    monkeypatch.setattr(J, "KILL_GRACE_S", 0.6)
    stubborn = ("import signal, time, sys\n"
                "signal.signal(signal.SIGTERM, lambda *a: print('ignoring SIGTERM', flush=True))\n"
                "print('ready', flush=True)\n"
                "time.sleep(60)\n")
    polite = "import time\nprint('ready', flush=True)\ntime.sleep(60)\n"
    spec = M.JobSpec("job-x", M.Executor.aggregate(M.AGGREGATE_EXECUTOR_NAME, M.AggregateExecutorConfig(
        updates=M.Receive.from_peers(["w"]), results=M.Send.to_peers(["w"]), optimizer=M.Nesterov(0.7, 0.9))))

    async def run(code):
        net = Network("w")
        logs = []
        ex = await J.ProcessExecutor(Connector(net), net, str(tmp_path), sys.executable, ["-c", code],
                                     log=logs.append).execute(spec, "t", "sched")
        for _ in range(200):
            if any("ready" in l for l in logs):
                break
            await asyncio.sleep(0.02)
        t0 = time.perf_counter()
        ex.cancel()
        status = await asyncio.wait_for(ex.wait(), 10)
        dt = time.perf_counter() - t0
        await net.close()
        return status, dt, logs, ex.work_dir

    status, dt, logs, wd = asyncio.run(run(stubborn))
    assert status == M.JobStatus.FAILED and 0.55 <= dt < 5 and any("ignoring SIGTERM" in l for l in logs)
    assert not os.path.exists(wd)
    monkeypatch.setattr(J, "KILL_GRACE_S", 5.0)
    status, dt, logs, wd = asyncio.run(run(polite))
    assert dt < 3.0 and not os.path.exists(wd)                 # SIGTERM was enough: the grace period was not spent


def test_binary_release_roundtrip(tmp_path):
    """`tools/make_release.py` -> `install.sh --from-release`: the artefact carries the prebuilt extensions (no
    sources of the native code), the installer verifies sha256 + manifest, unpacks and links launchers that run
    WITHOUT the checkout on the path; a tampered archive is refused."""
    import json
    import os
    import subprocess
    import sys
    import tarfile

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "make_release.py"), "--out", str(tmp_path / "dist"),
                          "--skip-cuda"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    info = json.loads(out.stdout.strip().splitlines()[-1])
    names = tarfile.open(info["artefact"]).getnames()
    assert any(n.endswith("MANIFEST.json") for n in names) and any("_core" in n and n.endswith(".so") for n in names)
    assert not any("/csrc/" in n for n in names)
    home = tmp_path / "home"
    env = dict(os.environ, HOME=str(home), HYPHA_NO_MODIFY_PATH="1", HYPHA_PYTHON=sys.executable)
    env.pop("PYTHONPATH", None)
    r = subprocess.run(["sh", os.path.join(root, "install.sh"), "--from-release", info["artefact"]], env=env,
                       capture_output=True, text=True, timeout=300, cwd=str(tmp_path))
    assert r.returncode == 0 and "sha256 verified" in r.stdout, r.stdout + r.stderr
    gw = home / ".local" / "bin" / "hypha-gateway"
    r = subprocess.run([str(gw), "init", "-o", str(tmp_path / "gw.toml")], env=env, capture_output=True, text=True,
                       timeout=120, cwd=str(tmp_path))
    assert r.returncode == 0 and (tmp_path / "gw.toml").exists(), r.stderr
    bad = tmp_path / "bad.tar.gz"
    blob = bytearray(open(info["artefact"], "rb").read())
    blob[len(blob) // 2] ^= 0xFF
    bad.write_bytes(bytes(blob))
    open(str(bad) + ".sha256", "w").write(open(info["artefact"] + ".sha256").read())
    r = subprocess.run(["sh", os.path.join(root, "install.sh"), "--from-release", str(bad)], env=env, capture_output=True,
                       text=True, timeout=120, cwd=str(tmp_path))
    assert r.returncode != 0 and "sha256 mismatch" in (r.stdout + r.stderr)


def test_hypha_data_prepare_cuts_a_corpus_into_slices(tmp_path, capsys):
    """`hypha-data prepare`: a flat token stream (.bin / .npy), packed samples or text become the slice directory a
    data node serves -- rows of `seq_len` ids, `samples_per_slice` per file, labels = inputs, ids range-checked."""
    import json

    import numpy as np
    import torch
    from safetensors.torch import load_file

    from hypha_b200.cli import main
    from hypha_b200.data import dataset_files
    from hypha_b200.executor.dataset import IterableStreamDataSet

    stream = np.arange(10 * 16 + 5, dtype=np.uint16) % 50
    src = tmp_path / "corpus.bin"
    stream.tofile(src)
    out = tmp_path / "toyset"
    assert main("data", ["prepare", str(src), "-o", str(out), "--seq-len", "16", "--samples-per-slice", "4",
                         "--vocab-size", "50"]) == 0
    summary = json.loads(capsys.readouterr().out)
    assert summary == {"dataset": "toyset", "num_slices": 3, "samples": 10, "seq_len": 16, "samples_per_slice": 4,
                       "tokens": 160, "max_token_id": 49, "inputs": ["input_ids", "labels"]}
    assert json.load(open(str(out) + ".json")) == summary
    files = dataset_files(str(out))
    assert [os.path.basename(f) for f in files] == ["slice-00000.safetensors", "slice-00001.safetensors",
                                                    "slice-00002.safetensors"]
    rows = list(IterableStreamDataSet(iter(files), ["input_ids", "labels"]))
    assert len(rows) == 10 and all(r["input_ids"].shape == (16,) and torch.equal(r["input_ids"], r["labels"]) for r in rows)
    assert torch.equal(torch.cat([r["input_ids"] for r in rows]), torch.from_numpy(stream[:160].astype(np.int64)))
    # refusals: a directory that already holds files, ids outside the vocabulary, a corpus shorter than one sample
    assert main("data", ["prepare", str(src), "-o", str(out), "--seq-len", "16"]) == 2
    assert main("data", ["prepare", str(src), "-o", str(tmp_path / "v"), "--seq-len", "16", "--vocab-size", "10"]) == 2
    assert main("data", ["prepare", str(src), "-o", str(tmp_path / "s"), "--seq-len", "4096"]) == 2
    err = capsys.readouterr().err
    assert "not empty" in err and "outside the vocabulary" in err and "shorter than one sample" in err
    # packed .npy samples, shuffled deterministically, last short slice dropped, no labels
    packed = np.arange(9 * 8, dtype=np.int64).reshape(9, 8)
    np.save(tmp_path / "packed.npy", packed)
    out2 = tmp_path / "packedset"
    assert main("data", ["prepare", str(tmp_path / "packed.npy"), "-o", str(out2), "--seq-len", "8",
                         "--samples-per-slice", "4", "--shuffle-seed", "1", "--no-labels", "--drop-last"]) == 0
    s2 = json.loads(capsys.readouterr().out)
    assert s2["num_slices"] == 2 and s2["samples"] == 8 and s2["inputs"] == ["input_ids"]
    got = torch.cat([load_file(f)["input_ids"] for f in dataset_files(str(out2))])
    assert sorted(got[:, 0].tolist()) != got[:, 0].tolist() and set(got[:, 0].tolist()) <= set(packed[:, 0].tolist())
    assert main("data", ["prepare", str(tmp_path / "packed.npy"), "-o", str(tmp_path / "w"), "--seq-len", "16"]) == 2
    # text needs a tokenizer (a local directory: nothing is downloaded); documents are joined by the EOS token
    (tmp_path / "t.txt").write_text("hello world\nworld again hello\n")
    assert main("data", ["prepare", str(tmp_path / "t.txt"), "-o", str(tmp_path / "t"), "--seq-len", "4"]) == 2
    capsys.readouterr()
    from tokenizers import Tokenizer, models, pre_tokenizers
    from transformers import PreTrainedTokenizerFast

    tk = Tokenizer(models.WordLevel({"<eos>": 0, "hello": 1, "world": 2, "again": 3, "[UNK]": 4}, unk_token="[UNK]"))
    tk.pre_tokenizer = pre_tokenizers.Whitespace()
    PreTrainedTokenizerFast(tokenizer_object=tk, eos_token="<eos>", unk_token="[UNK]").save_pretrained(str(tmp_path / "tok"))
    (tmp_path / "t.jsonl").write_text('{"text": "hello world"}\n\n{"text": "world again hello"}\n')
    for name in ("t.txt", "t.jsonl"):
        dst = tmp_path / ("text-" + name.split(".")[1])
        assert main("data", ["prepare", str(tmp_path / name), "-o", str(dst), "--seq-len", "7", "--tokenizer",
                             str(tmp_path / "tok")]) == 0
        assert json.loads(capsys.readouterr().out)["samples"] == 1
        assert load_file(dataset_files(str(dst))[0])["input_ids"].tolist() == [[1, 2, 0, 2, 3, 1, 0]]
