"""Byte counters per protocol: ``hypha.bandwidth.{inbound,outbound}.bytes``.

Reference: crates/telemetry/src/bandwidth.rs:44-66 wraps every libp2p stream and adds the bytes
read/written to two OTel counters with a ``protocols`` attribute.  Here every link of the
control-plane fabric and every peer-memory operation reports into the same structure."""
from __future__ import annotations

import threading
from typing import Dict

INBOUND = "hypha.bandwidth.inbound.bytes"
OUTBOUND = "hypha.bandwidth.outbound.bytes"


class BandwidthCounters:
    def __init__(self) -> None:
        self._lock = threading.Lock()
        self.inbound: Dict[str, int] = {}
        self.outbound: Dict[str, int] = {}

    def add(self, protocol: str, inbound: int = 0, outbound: int = 0) -> None:
        with self._lock:
            if inbound:
                self.inbound[protocol] = self.inbound.get(protocol, 0) + int(inbound)
            if outbound:
                self.outbound[protocol] = self.outbound.get(protocol, 0) + int(outbound)

    def total(self) -> Dict[str, int]:
        with self._lock:
            return {INBOUND: sum(self.inbound.values()), OUTBOUND: sum(self.outbound.values())}

    def snapshot(self) -> Dict[str, Dict[str, int]]:
        with self._lock:
            return {INBOUND: dict(self.inbound), OUTBOUND: dict(self.outbound)}
