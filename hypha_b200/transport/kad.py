"""Kademlia routing table of the fabric's DHT (the reference: crates/network/src/kad.rs -- identify results feed the
routing table through a CIDR filter `:394-432`, records and providers live on the `k` closest servers, lookups walk
towards the key `:482-737`; reserved address ranges crates/network/src/lib.rs:57-98).

* ids: ``sha256(peer id)`` / ``sha256(key)``, distance = XOR, 256 buckets by the length of the common prefix with our
  own id, `K` entries per bucket in least-recently-seen order; a full bucket keeps its old, still connected entries and
  replaces a stale one (Kademlia's preference for long-lived nodes) -- implemented in the C++ core
  (`csrc/core/kad.hpp`, bound as `_core.KadTable`), like the reference's table is native;
* only peers that announce themselves as DHT servers are routable, and only with addresses that pass the CIDR filter:
  in strict mode an address inside a reserved range (loopback, RFC 1918, link-local, CGNAT, multicast, documentation
  nets, ...) is not learnt, and a peer with no acceptable address does not enter the table;
* `closest(key, n)`: the n known servers nearest to the key -- where `put_record` stores, where `get_record` starts.
"""
from __future__ import annotations

import hashlib
import ipaddress
from typing import Iterable, List, Optional, Sequence, Tuple

K = 20                  # bucket size
ALPHA = 3               # parallelism of an iterative lookup
REPLICATION = 3         # servers a record is stored on

RESERVED_CIDRS: Tuple[str, ...] = (
    "0.0.0.0/8", "10.0.0.0/8", "100.64.0.0/10", "127.0.0.0/8", "169.254.0.0/16", "172.16.0.0/12", "192.0.0.0/24",
    "192.0.2.0/24", "192.168.0.0/16", "198.18.0.0/15", "198.51.100.0/24", "203.0.113.0/24", "224.0.0.0/4",
    "240.0.0.0/4", "::/128", "::1/128", "fc00::/7", "fe80::/10", "ff00::/8", "2001:db8::/32",
)


def key_id(s: str) -> int:
    return int.from_bytes(hashlib.sha256(s.encode()).digest(), "big")


def _host_of(addr: str) -> Optional[str]:
    for scheme in ("tcp://", "udp://"):
        if addr.startswith(scheme):
            host = addr[len(scheme):].rsplit(":", 1)[0]
            return host.strip("[]")
    return None


class CidrFilter:
    """Which advertised addresses may enter the routing table.  ``exclude``: ranges never learnt (strict mode passes
    `RESERVED_CIDRS`); ``allow``: exceptions inside them (a site's private range, the reference's `exclude_cidr`
    override).  memory:// and unix:// addresses are process-local and always pass."""

    def __init__(self, exclude: Sequence[str] = (), allow: Sequence[str] = ()):
        self.exclude = [ipaddress.ip_network(c) for c in exclude]
        self.allow = [ipaddress.ip_network(c) for c in allow]

    def accepts(self, addr: str) -> bool:
        host = _host_of(addr)
        if host is None:
            return addr.startswith(("memory://", "unix://"))
        try:
            ip = ipaddress.ip_address(host)
        except ValueError:
            return True                              # a DNS name: resolved at dial time
        if any(ip.version == n.version and ip in n for n in self.allow):
            return True
        return not any(ip.version == n.version and ip in n for n in self.exclude)

    def filter(self, addrs: Iterable[str]) -> List[str]:
        return [a for a in addrs if self.accepts(a)]


class RoutingTable:
    """Policy layer over the native table (`csrc/core/kad.hpp`: sha256 ids, XOR metric, k-buckets, closest): this
    class only decides WHICH addresses of a peer may be learnt (CIDR filter) and counts what it turned away."""

    def __init__(self, self_id: str, k: int = K, cidr: Optional[CidrFilter] = None):
        from ..ops.native import load_core

        self.self_id, self.k = self_id, k
        self.cidr = cidr or CidrFilter()
        self._t = load_core().KadTable(self_id, k)
        self.rejected_by_cidr = 0

    @property
    def buckets(self) -> List[List[str]]:
        """Peers per bucket (index = length of the id prefix shared with this node), least recently seen first."""
        return [self._t.bucket(i) for i in range(256)]

    def _bucket(self, peer: str) -> List[str]:
        return self._t.bucket(self._t.bucket_index(peer))

    def add(self, peer: str, addrs: Sequence[str], connected: bool = True) -> bool:
        """Learn (or refresh) a DHT server.  False when the CIDR filter leaves it without an address or its bucket is
        full of live entries."""
        if peer == self.self_id:
            return False
        ok = self.cidr.filter(addrs)
        if addrs and not ok:
            self.rejected_by_cidr += 1
            return False
        return bool(self._t.add(peer, list(ok), connected))

    def mark_disconnected(self, peer: str) -> None:
        self._t.mark_disconnected(peer)

    def remove(self, peer: str) -> None:
        self._t.remove(peer)

    def addrs(self, peer: str) -> List[str]:
        return list(self._t.addrs(peer))

    def peers(self) -> List[str]:
        return list(self._t.peers())

    def __len__(self) -> int:
        return len(self._t)

    def __contains__(self, peer: str) -> bool:
        return bool(self._t.contains(peer))

    def closest(self, key: str, n: int = K, include_self: bool = False) -> List[str]:
        return list(self._t.closest(key, n, include_self))
