"""One-process loopback cluster: gateway + data node + N train workers + 1 aggregate worker +
scheduler over ``memory://`` links (train executors are real subprocesses behind UDS bridges).

This is the north-star control-plane layout ("scheduler/gateway/leases/messages run as in-process
Python over a loopback transport") and what BASELINE config 1 (GPT-2 DiLoCo on CPU) runs on."""
from __future__ import annotations

import asyncio
import os
from typing import Any, Callable, Dict, List, Optional, Tuple

from . import config as C
from . import roles
from .scheduler import MemoryConnector

TRAIN_EXEC = {"class": "train", "name": "diloco-transformer", "runtime": "process", "cmd": "python",
              "args": ["-m", "hypha_b200.executor.training", "--socket", "{SOCKET_PATH}", "--work-dir", "{WORK_DIR}",
                       "--job", "{JOB_JSON}"]}
PS_EXEC = {"class": "aggregate", "name": "parameter-server", "runtime": "parameter-server"}


def worker_config(gateway: str, work_dir: str, executors: List[dict], gpu: float = 4.0,
                  price: float = 1.0, device: Optional[int] = None) -> C.WorkerConfig:
    return C.builder(C.WorkerConfig).with_overrides({
        "listen_addresses": [], "gateway_addresses": [gateway], "exclude_cidr": [], "work_dir": work_dir,
        "device": device,
        "resources": {"gpu": gpu, "cpu": 4.0, "memory": 8.0, "storage": 10.0},
        "offer": {"price": price, "floor": 0.5, "strategy": {"type": "whole"}}, "executors": executors}).build()


async def run_loopback_job(tmp: str, *, repo: str, write_model_config: Callable[[str], None],
                           write_dataset: Callable[[str], None], job: Dict[str, Any], num_workers: int = 2,
                           timeout_s: float = 600.0, log: Optional[Callable[[str], None]] = None,
                           omp_threads: int = 2, chaos: Optional[Callable[[Dict[str, Any], List[str]], Any]] = None,
                           spare_workers: int = 0, executor_env: Optional[Dict[str, Dict[str, str]]] = None
                           ) -> Tuple[Any, MemoryConnector, List[str], List[Any]]:
    """Returns ``(JobOutcome, metrics sink, executor logs, nodes)``.  ``chaos(nodes_by_peer_id, logs)`` is an
    optional coroutine run next to the job (fault injection: e.g. stop a worker node mid-run)."""
    registry = os.path.join(tmp, "registry")
    os.makedirs(os.path.join(registry, repo), exist_ok=True)
    write_model_config(os.path.join(registry, repo, "config.json"))
    os.environ["HYPHA_HF_LOCAL"] = registry
    os.environ["OMP_NUM_THREADS"] = str(omp_threads)
    # several executors share this machine's cores: OpenMP workers that spin while they wait burn the cores the
    # other executors need (measured: 15x the CPU time for the same job on a 3-core box)
    os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")
    os.environ.setdefault("GOMP_SPINCOUNT", "0")
    os.environ.setdefault("KMP_BLOCKTIME", "0")
    ds_dir = os.path.join(tmp, job["dataset"]["dataset"])
    write_dataset(ds_dir)
    gw_addr = f"memory://gw-{os.getpid()}-{abs(hash(tmp)) % 10 ** 8}"
    loop = asyncio.get_event_loop()
    logs: List[str] = []
    sink_log = log or logs.append

    def _log(m: str) -> None:
        logs.append(m)
        if log is not None:
            log(m)

    collective = (job.get("aggregation") or {}).get("mode") == "collective"
    # spare workers ask a higher price, so the allocator only takes them when a cheaper one is gone
    spares = [f"worker-{chr(ord('a') + num_workers + i)}" for i in range(spare_workers)]
    names = [f"worker-{chr(ord('a') + i)}" for i in range(num_workers)] + spares + ([] if collective else ["worker-ps"])
    started = [loop.create_future() for _ in range(2 + len(names))]
    gw_cfg = C.builder(C.GatewayConfig).with_overrides({"listen_addresses": [gw_addr]}).build()
    bg = [asyncio.ensure_future(roles.run_gateway(gw_cfg, peer_id="gateway", started=started[0]))]
    await started[0]
    data_cfg = C.builder(C.DataConfig).with_overrides({"listen_addresses": [], "gateway_addresses": [gw_addr],
                                                       "exclude_cidr": [], "dataset_path": ds_dir}).build()
    bg.append(asyncio.ensure_future(roles.run_data(data_cfg, peer_id="data", started=started[1])))
    for i, name in enumerate(names):
        wd = os.path.join(tmp, name)
        os.makedirs(wd, exist_ok=True)
        ex = [PS_EXEC] if name == "worker-ps" else [dict(TRAIN_EXEC, env=(executor_env or {}).get(name, {}))]
        price = 1.5 if name in spares else 1.0
        # train worker i drives GPU i (ignored on CPU; collective jobs pick the device from their rank)
        dev = i if name != "worker-ps" else None
        bg.append(asyncio.ensure_future(roles.run_worker(worker_config(gw_addr, wd, ex, price=price, device=dev), peer_id=name,
                                                         started=started[2 + i], log=_log)))
    nodes = [await f for f in started[1:]]
    over = {"listen_addresses": [], "gateway_addresses": [gw_addr], "exclude_cidr": [], "scheduler": {"job": job}}
    sched_cfg = C.builder(C.SchedulerConfig).with_overrides(over).build()
    sink = MemoryConnector()
    if chaos is not None:
        bg.append(asyncio.ensure_future(chaos({n.peer_id: n for n in nodes}, logs)))
    try:
        outcome = await asyncio.wait_for(
            roles.run_scheduler(sched_cfg, peer_id="scheduler", metrics_connector=sink, allocation_pause_s=0.7),
            timeout_s)
    finally:
        for n in nodes + [started[0].result()]:
            n.stop.set()
        await asyncio.wait(bg, timeout=20)
    return outcome, sink, logs, nodes
