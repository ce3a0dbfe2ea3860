"""Stream windowing + CIDR helpers (reference: crates/network/src/utils.rs:18-42,50-141,
reserved ranges crates/network/src/lib.rs:57-98) and identity derivation from a node key
(crates/network/src/cert.rs:30-80)."""
from __future__ import annotations

import asyncio
import hashlib
import ipaddress
import time
from typing import Any, AsyncIterator, Iterable, List, Optional

from ..config import DEFAULT_EXCLUDE_CIDR

RESERVED_IPV4 = [c for c in DEFAULT_EXCLUDE_CIDR if ":" not in c]
RESERVED_IPV6 = [c for c in DEFAULT_EXCLUDE_CIDR if ":" in c]


def parse_cidrs(cidrs: Iterable[str]) -> List[Any]:
    return [ipaddress.ip_network(c, strict=False) for c in cidrs]


def find_containing_cidr(addr: str, cidrs: Iterable[str]) -> Optional[str]:
    """The first CIDR that contains ``addr`` (an IP literal), else ``None``."""
    try:
        ip = ipaddress.ip_address(addr)
    except ValueError:
        return None
    for n in parse_cidrs(cidrs):
        if ip.version == n.version and ip in n:
            return str(n)
    return None


def is_excluded(addr: str, cidrs: Iterable[str]) -> bool:
    return find_containing_cidr(addr, cidrs) is not None


class Batched:
    """``async for batch in Batched(stream, limit, window_s)``: yields lists of up to ``limit``
    items, flushing a partial batch once ``window_s`` elapsed since its first item.  The worker
    arbiter uses limit 100 / 200 ms (arbiter.rs:25-26)."""

    def __init__(self, source: AsyncIterator[Any], limit: int, window_s: float):
        self.source, self.limit, self.window_s = source, limit, window_s

    def __aiter__(self):
        return self._gen()

    async def _gen(self):
        it = self.source.__aiter__()
        batch: List[Any] = []
        deadline = None
        nxt: Optional[asyncio.Future] = None
        while True:
            if nxt is None:
                nxt = asyncio.ensure_future(it.__anext__())
            timeout = None if deadline is None else max(0.0, deadline - time.monotonic())
            done, _ = await asyncio.wait({nxt}, timeout=timeout)
            if nxt in done:
                try:
                    item = nxt.result()
                except StopAsyncIteration:
                    if batch:
                        yield batch
                    return
                nxt = None
                if not batch:
                    deadline = time.monotonic() + self.window_s
                batch.append(item)
                if len(batch) >= self.limit:
                    yield batch
                    batch, deadline = [], None
            elif batch:
                yield batch
                batch, deadline = [], None


def peer_id_from_public_bytes(pub: bytes) -> str:
    """The peer id of a raw Ed25519 public key (one scheme for private keys and certificates)."""
    import base64

    return "12D3Koo" + base64.b32encode(hashlib.sha256(pub).digest()).decode().rstrip("=")[:40]


def peer_id_from_cert_der(der: bytes) -> str:
    """Peer id bound to a TLS peer certificate: derived from the certificate's subject public key, so a node
    can only ever speak under the id of the key its CA-issued certificate was made for (the reference binds
    the libp2p PeerId to the certificate key the same way, crates/network/src/cert.rs:30-80)."""
    from cryptography import x509
    from cryptography.hazmat.primitives import serialization

    pub = x509.load_der_x509_certificate(der).public_key()
    try:
        raw = pub.public_bytes(serialization.Encoding.Raw, serialization.PublicFormat.Raw)
    except ValueError:       # not an Ed25519/X25519-style key: fall back to the DER SubjectPublicKeyInfo
        raw = pub.public_bytes(serialization.Encoding.DER, serialization.PublicFormat.SubjectPublicKeyInfo)
    return peer_id_from_public_bytes(raw)


def peer_id_from_key_pem(path: str) -> str:
    """Stable peer id derived from the node's Ed25519 public key (``12D3Koo``-style prefix,
    base32 of sha256(pubkey)); the reference derives the libp2p PeerId the same way from the
    PKCS#8 key loaded at start-up."""
    import base64

    from cryptography.hazmat.primitives import serialization

    key = serialization.load_pem_private_key(open(path, "rb").read(), password=None)
    pub = key.public_key().public_bytes(serialization.Encoding.Raw, serialization.PublicFormat.Raw)
    return "12D3Koo" + base64.b32encode(hashlib.sha256(pub).digest()).decode().rstrip("=")[:40]
