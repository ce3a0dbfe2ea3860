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
REFRESH_BUCKETS = 8     # buckets refreshed by a lookup of their own (the 8 far ones hold 255/256 of all ids)

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
        self._refresh: List[str] = []

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

    def refresh_keys(self, buckets: int = REFRESH_BUCKETS) -> List[str]:
        """One lookup key per bucket ``0 .. buckets-1``: a string whose hashed id shares exactly ``i`` leading bits with
        this node's id, i.e. falls into bucket ``i``'s range.  Kademlia refreshes a bucket by looking up a random id
        inside it; ids here are hashes of strings, so the string is found by trial (2^(i+1) hashes on average -- the
        far buckets, which hold almost all of a network, cost a handful; the near ones are covered by the lookup of
        the node's own id).  Computed once per node."""
        if len(self._refresh) < buckets:
            me = key_id(self.self_id)
            want = set(range(len(self._refresh), buckets))
            found: dict = {}
            n = 0
            while want and n < (1 << (buckets + 4)):
                s = f"hypha-kad-refresh/{self.self_id}/{n}"
                n += 1
                x = key_id(s) ^ me
                i = 256 - x.bit_length()               # length of the common prefix
                if i in want:
                    want.discard(i)
                    found[i] = s
            for i in range(len(self._refresh), buckets):
                if i not in found:
                    break
                self._refresh.append(found[i])
        return list(self._refresh[:buckets])


# ---------------------------------------------------------------------------------------------------------------------
# record store of a DHT node
# ---------------------------------------------------------------------------------------------------------------------
RECORD_TTL_S = 3600.0           # a record nobody republishes disappears (publishers republish every few minutes)
MAX_RECORD_VALUE = 64 * 1024    # bytes; the reference's records are a few dozen bytes (DataRecord{num_slices})
MAX_RECORD_KEY = 256
MAX_RECORDS = 16384             # per node
MAX_RECORDS_PER_PUBLISHER = 1024
MAX_PROVIDERS_PER_KEY = 64


class StoredRecord(tuple):
    """``(value, publisher)`` -- what `get_record` hands out -- with the bookkeeping as attributes."""

    def __new__(cls, value: bytes, publisher: str, seq: int = 0, sig: bytes = b"", pub: bytes = b"",
                expires: float = float("inf")):
        self = super().__new__(cls, (value, publisher))
        self.seq, self.sig, self.pub, self.expires = int(seq), bytes(sig or b""), bytes(pub or b""), expires
        return self

    value = property(lambda self: self[0])
    publisher = property(lambda self: self[1])


class RecordStore:
    """Records and provider sets a node keeps for the DHT, with the limits a node reachable by strangers needs:

    * size: values <= 64 KiB, keys <= 256 bytes, <= 16384 records per node and <= 1024 per publisher -- a peer cannot
      grow the store without bound; a full store evicts expired entries first and then refuses;
    * lifetime: an entry lives `RECORD_TTL_S` after it was last stored (the publisher's upkeep timer stores it again
      every few minutes), so what a departed peer left behind goes away;
    * ownership (``enforce_owner``, on under a PKI): while a key's record is alive only its publisher can replace it
      -- another certificate holder cannot take over a dataset's key and point the schedulers at its own data node;
    * order: a publisher's records carry a sequence number (its clock in ms) under its signature; an older copy --
      a replay by a server that kept one -- does not roll a newer one back.

    Reads like a dict of ``key -> (value, publisher)``."""

    def __init__(self, ttl_s: float = RECORD_TTL_S, enforce_owner: bool = False, clock=None):
        import time

        self.ttl_s, self.enforce_owner = ttl_s, enforce_owner
        self._clock = clock or time.monotonic
        self._recs: dict = {}
        self._per_pub: dict = {}
        self._provs: dict = {}                # key -> {provider: expires}
        self.refused: dict = {}               # reason -> count

    # -- records ------------------------------------------------------------------------------------------------------
    def _no(self, why: str) -> str:
        self.refused[why] = self.refused.get(why, 0) + 1
        return why

    def _drop(self, key: str) -> None:
        rec = self._recs.pop(key, None)
        if rec is not None:
            n = self._per_pub.get(rec.publisher, 0) - 1
            if n > 0:
                self._per_pub[rec.publisher] = n
            else:
                self._per_pub.pop(rec.publisher, None)

    def expire(self) -> int:
        now, n = self._clock(), 0
        for key in [k for k, r in self._recs.items() if r.expires <= now]:
            self._drop(key)
            n += 1
        for key in list(self._provs):
            live = {p: e for p, e in self._provs[key].items() if e > now}
            n += len(self._provs[key]) - len(live)
            if live:
                self._provs[key] = live
            else:
                del self._provs[key]
        return n

    def put(self, key: str, value: bytes, publisher: str, seq: int = 0, sig: bytes = b"", pub: bytes = b"",
            local: bool = False) -> str:
        """"stored", or why not ("too-large" | "owned" | "stale" | "full").  ``local``: this node's own record (no
        quota, never expires while the node republishes it)."""
        if not isinstance(key, str) or not isinstance(publisher, str) or len(key.encode()) > MAX_RECORD_KEY \
                or len(value) > MAX_RECORD_VALUE or len(publisher) > MAX_RECORD_KEY:
            return self._no("too-large")
        now = self._clock()
        cur = self._recs.get(key)
        if cur is not None and cur.expires <= now:
            self._drop(key)
            cur = None
        if cur is not None:
            if cur.publisher != publisher:
                if self.enforce_owner and not local:
                    return self._no("owned")
            elif int(seq) < cur.seq:
                return self._no("stale")
        if not local and (cur is None or cur.publisher != publisher):
            if self._per_pub.get(publisher, 0) >= MAX_RECORDS_PER_PUBLISHER or len(self._recs) >= MAX_RECORDS:
                self.expire()
                if self._per_pub.get(publisher, 0) >= MAX_RECORDS_PER_PUBLISHER or len(self._recs) >= MAX_RECORDS:
                    return self._no("full")
        if cur is not None:
            self._drop(key)
        self._recs[key] = StoredRecord(bytes(value), publisher, seq, sig, pub,
                                       float("inf") if local else now + self.ttl_s)
        self._per_pub[publisher] = self._per_pub.get(publisher, 0) + 1
        return "stored"

    def get(self, key: str, default=None):
        rec = self._recs.get(key)
        if rec is None:
            return default
        if rec.expires <= self._clock():
            self._drop(key)
            return default
        return rec

    def __getitem__(self, key: str):
        rec = self.get(key)
        if rec is None:
            raise KeyError(key)
        return rec

    def __setitem__(self, key: str, vp) -> None:
        self._drop(key)
        self._recs[key] = StoredRecord(vp[0], vp[1], expires=self._clock() + self.ttl_s)
        self._per_pub[vp[1]] = self._per_pub.get(vp[1], 0) + 1

    def __contains__(self, key) -> bool:
        return self.get(key) is not None

    def __len__(self) -> int:
        return len(self._recs)

    def keys(self):
        return list(self._recs)

    def clear(self) -> None:
        self._recs.clear()
        self._per_pub.clear()
        self._provs.clear()

    # -- providers ----------------------------------------------------------------------------------------------------
    def add_provider(self, key: str, provider: str) -> str:
        if not isinstance(key, str) or len(key.encode()) > MAX_RECORD_KEY or len(provider) > MAX_RECORD_KEY:
            return self._no("too-large")
        now = self._clock()
        cur = self._provs.get(key)
        if cur is None:
            if len(self._provs) >= MAX_RECORDS:
                self.expire()
                if len(self._provs) >= MAX_RECORDS:
                    return self._no("full")
            cur = self._provs[key] = {}
        if provider not in cur and len(cur) >= MAX_PROVIDERS_PER_KEY:
            for p in [p for p, e in cur.items() if e <= now]:
                del cur[p]
            if len(cur) >= MAX_PROVIDERS_PER_KEY:
                return self._no("full")
        cur[provider] = now + self.ttl_s
        return "stored"

    def providers(self, key: str) -> set:
        now = self._clock()
        return {p for p, e in self._provs.get(key, {}).items() if e > now}
