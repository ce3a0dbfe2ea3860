"""The multi-box outer step (`DiLoCoOuter(backend="sharded")`): stock collectives of the process group with the fp32
outer state sharded over all ranks.  Three gloo processes against a single-process model of the arithmetic, for a
parameter count that leaves the last shard short and for one that leaves it EMPTY, plus checkpoint / resume."""
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("world,n", [(3, 10248), (3, 2048), (2, 4096)])
def test_sharded_outer_step_matches_the_model(world, n):
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "mp_sharded_outer_check.py"), str(n)],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    try:
        for p in procs:
            out, _ = p.communicate(timeout=300)
            outs.append(out)
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r}:\n{out[-3000:]}"
        assert "sharded outer step ok" in out


def test_auto_backend_picks_by_team_size():
    """One NVSwitch box (<= 8 ranks with peer memory) -> fused; anything larger must not ask for peer memory."""
    import inspect

    from hypha_b200 import trainer
    from hypha_b200.parallel import diloco

    src = inspect.getsource(trainer.DiLoCoTrainer.create)
    assert "group.world <= 8" in src
    assert "sharded" in inspect.getsource(diloco.DiLoCoOuter.__init__)


def test_rendezvous_of_later_generations_stays_on_the_configured_host():
    from hypha_b200.scheduler.job import rendezvous_series

    local = rendezvous_series("tcp://127.0.0.1:29400")
    a, b = local(), local()
    assert a.startswith("tcp://127.0.0.1:") and b.startswith("tcp://127.0.0.1:")
    far = rendezvous_series("tcp://node-a.cluster:29400")
    assert [far(), far()] == ["tcp://node-a.cluster:29401", "tcp://node-a.cluster:29402"]
    v6 = rendezvous_series("tcp://[fd00::1]:4000")
    assert v6() == "tcp://[fd00::1]:4001"
