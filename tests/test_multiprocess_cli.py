"""The deployment path of the quickstart: every role is its own OS process started through its CLI,
talking over tcp:// links with mutual TLS (certificates from hypha-certutil), the scheduler job given as
TOML.  The in-process loopback tests cover the logic; this one covers CLI + config layering + sockets +
TLS + relay through the gateway + process executors together."""
import json
import os
import signal
import socket
import subprocess
import sys
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _spawn(args, env, log):
    return subprocess.Popen([sys.executable, "-m"] + args, cwd=ROOT, env=env, stdout=log, stderr=subprocess.STDOUT,
                            start_new_session=True)


@pytest.mark.timeout(300)
@pytest.mark.parametrize("scheme", ["tcp", "udp"])
def test_roles_as_separate_processes_over_tcp_with_mtls(tmp_path, scheme):
    """``tcp``: TLS 1.3 links.  ``udp``: the same deployment over the encrypted datagram transport (the reference's
    QUIC listeners) -- every role listens on and dials udp:// addresses, certificates authenticate the handshake."""
    from hypha_b200 import certutil
    from hypha_b200.executor.dataset import write_synthetic_token_slices
    from hypha_b200.models.gpt2 import GPT2Config

    tmp = str(tmp_path)
    # PKI: root -> org -> one certificate per node
    rc, rk = certutil.make_root("Acme", out_dir=tmp)
    oc, ok = certutil.make_org(rc, rk, "Acme", name="acme-org", out_dir=tmp)
    tls = {}
    for n in ("gateway", "data", "worker-1", "worker-2", "worker-ps", "scheduler"):
        c, k, t = certutil.make_node(oc, ok, n, sans=("127.0.0.1",), out_dir=tmp)
        tls[n] = ["cert_pem", c, "key_pem", k, "trust_pem", t]
    registry = os.path.join(tmp, "registry", "local", "gpt2-tiny")
    os.makedirs(registry)
    GPT2Config.tiny().to_json(os.path.join(registry, "config.json"))
    ds = os.path.join(tmp, "toyset")
    write_synthetic_token_slices(ds, num_slices=6, samples_per_slice=16, seq_len=32, vocab_size=512)
    port = _free_port()
    gw = f"{scheme}://127.0.0.1:{port}"
    env = dict(os.environ, HYPHA_HF_LOCAL=os.path.join(tmp, "registry"), OMP_NUM_THREADS="2",
               HYPHA_EXECUTOR_DEVICE="cpu", PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))

    def toml(name, body):
        p = os.path.join(tmp, f"{name}.toml")
        t = tls[name]
        head = "".join(f'{t[i]} = "{t[i + 1]}"\n' for i in range(0, 6, 2)) + 'exclude_cidr = ["0.0.0.0/32"]\n'
        if "listen_addresses" not in body:
            head += f'listen_addresses = ["{scheme}://127.0.0.1:0"]\n'
        open(p, "w").write(head + body)
        return p

    procs, logs = [], {}

    def start(name, module, cfg_path, extra=()):
        logs[name] = open(os.path.join(tmp, f"{name}.log"), "w")
        p = _spawn([module, "run", "-c", cfg_path, "--peer-id", name, *extra], env, logs[name])
        procs.append(p)
        return p

    try:
        start("gateway", "hypha_b200.cli.gateway", toml("gateway", f'listen_addresses = ["{gw}"]\n'))
        deadline = time.time() + 30
        while True:                                  # probe (over TLS) until the gateway is healthy
            r = subprocess.run([sys.executable, "-c",
                                "import asyncio,sys; from hypha_b200 import roles; "
                                f"sys.exit(0 if asyncio.run(roles.probe('{gw}', 2.0, tls=dict(cert_pem='{tls['scheduler'][1]}', "
                                f"key_pem='{tls['scheduler'][3]}', trust_pem='{tls['scheduler'][5]}'))) else 1)"],
                               cwd=ROOT, env=env)
            if r.returncode == 0:
                break
            assert time.time() < deadline, open(os.path.join(tmp, "gateway.log")).read()[-2000:]
            time.sleep(0.5)
        common = f'gateway_addresses = ["{gw}"]\n'
        start("data", "hypha_b200.cli.data", toml("data", common + f'dataset_path = "{ds}"\n'))
        train = ('[[executors]]\nclass = "train"\nname = "diloco-transformer"\nruntime = "process"\ncmd = "python"\n'
                 'args = ["-m", "hypha_b200.executor.training", "--socket", "{SOCKET_PATH}", "--work-dir", "{WORK_DIR}", '
                 '"--job", "{JOB_JSON}"]\n')
        res = '[resources]\ngpu = 4.0\ncpu = 4.0\nmemory = 8.0\nstorage = 10.0\n[offer]\nprice = 1.0\nfloor = 0.5\n'
        for w in ("worker-1", "worker-2"):
            wd = os.path.join(tmp, w)
            os.makedirs(wd)
            start(w, "hypha_b200.cli.worker", toml(w, common + f'work_dir = "{wd}"\n' + res + train))
        wd = os.path.join(tmp, "worker-ps")
        os.makedirs(wd)
        start("worker-ps", "hypha_b200.cli.worker", toml("worker-ps", common + f'work_dir = "{wd}"\n' + res +
              '[[executors]]\nclass = "aggregate"\nname = "parameter-server"\nruntime = "parameter-server"\n'))
        time.sleep(2.0)
        job = ('[scheduler.job]\ntype = "diloco"\n[scheduler.job.model]\nrepository = "local/gpt2-tiny"\n'
               'filenames = ["config.json"]\ntype = "causal-lm"\ninput_names = ["input_ids", "labels"]\n'
               '[scheduler.job.preprocessor]\nrepository = ""\n'
               '[scheduler.job.dataset]\ndataset = "toyset"\n[scheduler.job.rounds]\navg_samples_between_updates = 40\n'
               'update_rounds = 2\nmax_batch_size = 4\n[scheduler.job.inner_optimizer]\nlearning-rate = 0.003\n'
               '[scheduler.job.resources]\nnum_workers = 2\n[scheduler.job.resources.worker]\ngpu = 1.0\ncpu = 1.0\n'
               'memory = 1.0\nstorage = 0.0\n[scheduler.job.resources.parameter_server]\ngpu = 0.0\ncpu = 1.0\n'
               'memory = 1.0\nstorage = 0.0\n[scheduler.job.resources.worker_price]\nbid = 1.0\nmax = 2.0\n'
               '[scheduler.job.resources.parameter_server_price]\nbid = 1.0\nmax = 2.0\n')
        sched_cfg = toml("scheduler", common + job)
        out = subprocess.run([sys.executable, "-m", "hypha_b200.cli.scheduler", "run", "-c", sched_cfg, "--peer-id",
                              "scheduler"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=200)
        tail = "\n".join(open(os.path.join(tmp, f"{n}.log")).read()[-1500:] for n in ("worker-1", "worker-ps"))
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:] + tail
        res_line = json.loads(out.stdout.strip().splitlines()[-1])
        assert res_line["finished"] and res_line["rounds"] == 2
        # under mTLS every node speaks under the peer id of its own key (the --peer-id names are ignored)
        from hypha_b200.transport.utils import peer_id_from_key_pem as pid

        ids = {n: pid(tls[n][3]) for n in tls}
        assert sorted(res_line["workers"]) == sorted([ids["worker-1"], ids["worker-2"]])
        assert res_line["parameter_server"] == ids["worker-ps"]
    finally:
        for p in procs:
            try:
                os.killpg(p.pid, signal.SIGTERM)
            except ProcessLookupError:
                pass
        for p in procs:
            try:
                p.wait(10)
            except subprocess.TimeoutExpired:
                os.killpg(p.pid, signal.SIGKILL)
        for f in logs.values():
            f.close()
