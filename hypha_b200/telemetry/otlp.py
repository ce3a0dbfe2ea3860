"""Env-style parsers for OTLP settings (reference: crates/telemetry/src/{otlp,attributes}.rs).

``Attributes`` / ``Headers`` use the W3C-baggage-like ``k=v,k=v`` syntax of
``OTEL_RESOURCE_ATTRIBUTES`` / ``OTEL_EXPORTER_OTLP_HEADERS`` (percent-decoded values);
``Protocol`` is ``grpc`` | ``http/protobuf``; ``Endpoint`` must be an http(s) URL."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional
from urllib.parse import quote, unquote, urlparse


class OtlpConfigError(ValueError):
    pass


def parse_kv_list(text: Optional[str]) -> Dict[str, str]:
    out: Dict[str, str] = {}
    if not text:
        return out
    for item in text.split(","):
        item = item.strip()
        if not item:
            continue
        if "=" not in item:
            raise OtlpConfigError(f"expected key=value, got {item!r}")
        k, v = item.split("=", 1)
        k = k.strip()
        if not k:
            raise OtlpConfigError(f"empty key in {item!r}")
        out[k] = unquote(v.strip())
    return out


def format_kv_list(d: Dict[str, str]) -> str:
    return ",".join(f"{k}={quote(str(v), safe='')}" for k, v in d.items())


@dataclass(frozen=True)
class Protocol:
    value: str

    @classmethod
    def parse(cls, text: Optional[str]) -> "Protocol":
        t = (text or "grpc").strip().lower()
        if t in ("grpc",):
            return cls("grpc")
        if t in ("http/protobuf", "http", "http-protobuf"):
            return cls("http/protobuf")
        raise OtlpConfigError(f"unsupported OTLP protocol {text!r}")


@dataclass(frozen=True)
class Endpoint:
    url: str

    @classmethod
    def parse(cls, text: Optional[str]) -> Optional["Endpoint"]:
        if not text:
            return None
        u = urlparse(text.strip())
        if u.scheme not in ("http", "https") or not u.netloc:
            raise OtlpConfigError(f"OTLP endpoint must be an http(s) URL, got {text!r}")
        return cls(text.strip().rstrip("/"))

    def signal(self, protocol: Protocol, name: str) -> str:
        """HTTP exporters post to ``<endpoint>/v1/<signal>`` (tracing.rs:38-103)."""
        return f"{self.url}/v1/{name}" if protocol.value == "http/protobuf" else self.url
