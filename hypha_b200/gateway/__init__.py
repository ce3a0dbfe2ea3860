"""hypha-gateway: bootstrap / relay / DHT-server node (reference: crates/gateway/src/bin/hypha-gateway.rs:36-164,
behaviours crates/gateway/src/network.rs:40-49,82-174).

The gateway runs no job logic.  What it is comes from the two flags it starts its fabric node with:

* ``relay_server``: frames between peers without a direct link travel through it, newcomers get its peer table
  (so they can dial each other directly) and the address swap of a direct-connection upgrade goes through it
  (`Network.upgrade_direct`);
* ``dht_server``: it enters other nodes' Kademlia routing tables, stores records / provider entries, replicates what
  it accepted from a publisher to the other closest servers and names closer servers to lookups it cannot answer.

Readiness (the `/hypha-health/0.0.1` answer and `hypha-gateway probe`) is "listening" (hypha-gateway.rs:78,121).
"""
from __future__ import annotations

import asyncio
from typing import Optional

from .. import config as C


async def run_gateway(cfg: C.GatewayConfig, *, peer_id: Optional[str] = None,
                      started: Optional[asyncio.Future] = None) -> None:
    from ..roles import start_node

    node = await start_node("gateway", cfg, peer_id=peer_id, relay_server=True, dht_server=True,
                            need_bootstrap=False)
    node.ready.set()                        # gateway readiness == listening (hypha-gateway.rs:78,121)
    node.telemetry.logger.info("gateway %s listening on %s", node.peer_id, node.network.listen_addrs)
    if started is not None and not started.done():
        started.set_result(node)
    try:
        await node.stop.wait()
    finally:
        await node.shutdown()
