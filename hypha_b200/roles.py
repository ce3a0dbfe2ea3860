"""The four long-running roles -- gateway, data node, worker, scheduler -- as asyncio programs.

Reference main flows: crates/gateway/src/bin/hypha-gateway.rs:36-164,
crates/data/src/bin/hypha-data.rs:33-233, crates/worker/src/bin/hypha-worker.rs:42-283,
crates/scheduler/src/bin/hypha-scheduler.rs:54-432.  Common start-up sequence for every role:
telemetry -> Network -> listen -> external addresses -> dial gateways -> DHT bootstrap ->
health handler -> role logic -> graceful shutdown.

Roles run either all in one process over ``memory://`` links (the north-star layout: the 8 GPUs
of a box are the workers) or as separate processes over ``unix://`` / ``tcp://``.
"""
from __future__ import annotations

import asyncio
import json
import os
from dataclasses import dataclass, field
from typing import Any, Callable, Optional

from . import config as C
from . import messages as M
from .scheduler import AllocatorError, Connector as MetricsConnector, TaskError
from .scheduler.job import DiLoCoJobRun, JobError, JobOutcome as _JobOutcome, worker_batch_size as _worker_batch_size
from .telemetry import Telemetry
from .transport import Network, NetworkError, RecordNotFound
from .worker import Arbiter, Connector, JobManager

ALLOCATION_PAUSE_S = 1.0          # hypha-scheduler.rs:241 (let losers' tentative leases expire)


GATEWAY_REDIAL_S = float(os.environ.get("HYPHA_GATEWAY_REDIAL_S", "3"))


class RoleError(Exception):
    pass


@dataclass
class Node:
    """Started network + telemetry of one role."""

    name: str
    network: Network
    telemetry: Telemetry
    ready: asyncio.Event = field(default_factory=asyncio.Event)
    stop: asyncio.Event = field(default_factory=asyncio.Event)
    gateways: dict = field(default_factory=dict)             # configured address -> peer id (None: not reached yet)
    on_gateway_reconnect: list = field(default_factory=list)  # async callbacks(addr, peer) after a successful re-dial
    background: list = field(default_factory=list)

    @property
    def peer_id(self) -> str:
        return self.network.peer_id

    async def maintain_gateways(self, interval_s: Optional[float] = None) -> None:
        """Keep the links to the configured gateways up: a gateway that restarts (or a link that drops) is
        re-dialed every ``interval_s``; after a successful re-dial the role's callbacks run (the data node
        re-announces its dataset: the DHT records lived in the old gateway process).  Topic subscriptions and
        addresses travel with the hello frame, so gossip works again right away."""
        interval_s = GATEWAY_REDIAL_S if interval_s is None else interval_s
        while not self.stop.is_set():
            try:
                await asyncio.wait_for(self.stop.wait(), interval_s)
                return
            except asyncio.TimeoutError:
                pass
            for addr, peer in list(self.gateways.items()):
                link = self.network.links.get(peer) if peer is not None else None
                if link is not None and not link.closed:
                    continue
                if link is not None:                      # in-memory links do not notice the other side's exit
                    await self.network.disconnect(peer)
                try:
                    new_peer = await self.network.dial(addr, timeout=min(5.0, max(1.0, interval_s)))
                except NetworkError:
                    continue
                self.gateways[addr] = new_peer
                self.telemetry.logger.info("%s re-connected to gateway %s (%s)", self.name, addr, new_peer)
                for cb in list(self.on_gateway_reconnect):
                    try:
                        await cb(addr, new_peer)
                    except Exception as exc:      # noqa: BLE001 - a failing callback must not stop the loop
                        self.telemetry.logger.warning("gateway reconnect callback failed: %s", exc)

    async def shutdown(self) -> None:
        self.stop.set()
        for t in self.background:
            t.cancel()
        await self.network.close()
        self.telemetry.shutdown()


async def start_node(role: str, cfg: Any, *, peer_id: Optional[str] = None, relay_server: bool = False,
                     dht_server: bool = False, need_bootstrap: bool = True) -> Node:
    tel = Telemetry.from_config(f"hypha-{role}", cfg)
    C.load_tls(cfg)                         # validates that the configured PEM files are readable
    tls = {k: getattr(cfg, k, "") for k in ("cert_pem", "key_pem", "trust_pem", "crls_pem")}
    if getattr(cfg, "key_pem", ""):
        from .transport.utils import peer_id_from_key_pem

        derived = peer_id_from_key_pem(cfg.key_pem)      # identity = public key, like the reference
        if peer_id is not None and peer_id != derived and any(
                str(a).startswith(("tcp://", "udp://")) for a in list(cfg.listen_addresses) + list(getattr(cfg, "gateway_addresses", []) or [])):
            # peers check the claimed id against the TLS certificate: a chosen name would be refused on every link
            tel.logger.warning("--peer-id %s ignored: with mTLS the peer id is the one of the node's key (%s)", peer_id, derived)
            peer_id = derived
        elif peer_id is None:
            peer_id = derived
    net = Network(peer_id or f"{role}-{M.new_id()[:8]}", agent=f"hypha-{role}", relay_server=relay_server,
                  dht_server=dht_server, tls=tls,         # certificates authenticate tcp:// (mTLS) and udp:// links
                  exclude_cidrs=tuple(getattr(cfg, "exclude_cidr", []) or []))   # kept out of the Kademlia routing table
    if any(a.startswith("tcp://") for a in cfg.listen_addresses) and net.tls is None:
        tel.logger.warning("tcp:// listener without cert_pem/key_pem/trust_pem: the link is NOT encrypted")
    if any(a.startswith("udp://") for a in cfg.listen_addresses) and net.tls is None:
        tel.logger.warning("udp:// listener without cert_pem/key_pem/trust_pem: the link is encrypted but peers are NOT authenticated")
    tel.register_bandwidth(net.counters)
    node = Node(role, net, tel)
    net.enable_dht_handlers()
    net.start_dht_maintenance(float(os.environ.get("HYPHA_DHT_REFRESH_S", "300")))
    for addr in cfg.listen_addresses:
        await net.listen(addr)
    for addr in getattr(cfg, "external_addresses", []) or []:
        net.add_external_address(addr)

    async def health(peer, msg):            # /hypha-health/0.0.1: healthy == ready
        return M.HealthResponse(node.ready.is_set())

    net.on(M.HEALTH_PROTOCOL, None, health)
    gateways = getattr(cfg, "gateway_addresses", []) or []
    dialed = 0
    for addr in gateways:
        try:
            node.gateways[addr] = await net.dial(addr)
            dialed += 1
        except NetworkError as exc:
            node.gateways[addr] = None
            tel.logger.warning("dial gateway %s failed: %s", addr, exc)
    if gateways and dialed == 0:
        raise RoleError(f"{role}: none of the gateways {gateways} could be dialed")
    if need_bootstrap and gateways:
        await net.wait_for_bootstrap(10.0)
    if gateways:
        node.background.append(asyncio.ensure_future(node.maintain_gateways()))
    return node


async def probe(addr: str, timeout: float = 5.0, tls: Optional[dict] = None) -> bool:
    """``probe <addr>``: dial, send ``health::Request``, true iff healthy
    (hypha-scheduler.rs:494-535)."""
    pid = f"probe-{M.new_id()[:8]}"
    if tls and tls.get("key_pem"):
        from .transport.utils import peer_id_from_key_pem

        pid = peer_id_from_key_pem(tls["key_pem"])       # under mTLS a node speaks under the id of its key only
    net = Network(pid, tls=tls)
    try:
        peer = await net.dial(addr, timeout=timeout)
        resp = await net.request(peer, M.HEALTH_PROTOCOL, M.HealthRequest(), timeout=timeout)
        return bool(M.HealthResponse.from_wire(resp).healthy)
    except NetworkError:
        return False
    finally:
        await net.close()


# ---------------------------------------------------------------------------------------------
# gateway and data node: `hypha_b200/gateway`, `hypha_b200/data` (re-exported here)
# ---------------------------------------------------------------------------------------------
from .data import dataset_files, run_data  # noqa: E402,F401
from .gateway import run_gateway  # noqa: E402,F401


# ---------------------------------------------------------------------------------------------
# worker
# ---------------------------------------------------------------------------------------------
async def run_worker(cfg: C.WorkerConfig, *, peer_id: Optional[str] = None,
                     started: Optional[asyncio.Future] = None, log: Callable[[str], None] = print) -> None:
    C.validate(cfg)
    node = await start_node("worker", cfg, peer_id=peer_id)
    net = node.network
    connector = Connector(net)
    exec_env = {"HYPHA_DEVICE": str(cfg.device)} if cfg.device is not None else {}
    jobs = JobManager(connector, net, cfg.work_dir, cfg.executors, log=log, executor_env=exec_env)
    cap = M.Resources(gpu=cfg.resources.gpu, cpu=cfg.resources.cpu, storage=cfg.resources.storage,
                      memory=cfg.resources.memory)
    arbiter = Arbiter(net, cap, jobs, price=cfg.offer.price, floor=cfg.offer.floor,
                      strategy=cfg.offer.strategy.type, log=log)
    node.arbiter = arbiter  # type: ignore[attr-defined]
    node.jobs = jobs  # type: ignore[attr-defined]
    from .telemetry.gpu_monitor import start_for_worker

    node.gpu_monitor = start_for_worker(cfg, node.telemetry, log)  # type: ignore[attr-defined]  # hypha.gpu.* gauges
    node.ready.set()                         # listening + bootstrapped (hypha-worker.rs:84-87,200)
    if started is not None and not started.done():
        started.set_result(node)
    try:
        await arbiter.run(node.stop)
    finally:
        if node.gpu_monitor is not None:  # type: ignore[attr-defined]
            node.gpu_monitor.stop()  # type: ignore[attr-defined]
        await jobs.shutdown()
        await node.shutdown()


# ---------------------------------------------------------------------------------------------
# scheduler
# ---------------------------------------------------------------------------------------------
JobOutcome = _JobOutcome          # re-exported: the scheduler role's result type
worker_batch_size = _worker_batch_size


async def get_data_provider(net: Network, dataset: str, retries: int = 20) -> tuple:
    for _ in range(retries):
        try:
            value, publisher = await net.get_record(dataset)
            return publisher, M.DataRecord.from_wire(json.loads(value))
        except RecordNotFound:
            await asyncio.sleep(0.25)
    raise RoleError(f"dataset {dataset!r} is not announced by any data node")


async def run_scheduler(cfg: C.SchedulerConfig, *, peer_id: Optional[str] = None,
                        started: Optional[asyncio.Future] = None,
                        metrics_connector: Optional[MetricsConnector] = None,
                        allocation_pause_s: float = ALLOCATION_PAUSE_S) -> JobOutcome:
    """hypha-scheduler: start the node, then allocate -> dispatch -> watch one job
    (`scheduler/job.py: DiLoCoJobRun`; reference flow hypha-scheduler.rs:54-432)."""
    C.validate(cfg)
    node = await start_node("scheduler", cfg, peer_id=peer_id)
    node.ready.set()
    if started is not None and not started.done():
        started.set_result(node)
    run = DiLoCoJobRun(node.network, cfg.scheduler.job, telemetry=node.telemetry, stop=node.stop,
                       data_provider=get_data_provider, metrics_connector=metrics_connector,
                       status_bridge=cfg.status_bridge, allocation_pause_s=allocation_pause_s)
    try:
        return await run.run()
    except (JobError, TaskError) as exc:      # incomplete allocation / a worker refused the dispatch
        raise RoleError(str(exc)) from exc
    except AllocatorError as exc:             # e.g. nobody answered the advertisement within the deadline
        raise RoleError(f"allocation failed: {exc}") from exc
    finally:
        await node.shutdown()
