"""Executor-side client of the job bridge (HTTP/1.1 over the job's Unix socket).

Public surface = what executors written for the reference expect (its client lives in
executors/accelerate/src/hypha/accelerate_executor/api.py): ``Session(socket_path)`` with ``fetch``,
``send_resource``, ``send_status`` and the ``receive`` context manager whose value iterates server-sent
events.  The implementation is organised around one request helper and an explicit SSE parser, adds
typed errors that carry the bridge's JSON error body, ``openapi()``, and keep-alive accounting on the
event stream (the bridge sends a comment line every 5 s, worker/bridge.py)."""
from __future__ import annotations

import json
from contextlib import contextmanager
from dataclasses import dataclass, field
from typing import Any, Dict, Iterable, Iterator, List, Optional

import httpx

_BASE = "http://hypha"                 # the host part is ignored on a Unix-socket transport
ROUTES = {"fetch": "/resources/fetch", "send": "/resources/send", "receive": "/resources/receive",
          "status": "/status/send", "openapi": "/openapi.json"}


class BridgeRequestError(httpx.HTTPStatusError):
    """Non-2xx answer of the bridge; ``detail`` is the ``error`` string of its JSON body when present."""

    def __init__(self, response: httpx.Response):
        try:
            self.detail = str(response.json().get("error", ""))
        except (ValueError, AttributeError):
            self.detail = response.text[:200]
        super().__init__(f"bridge answered {response.status_code} on {response.request.url.path}: {self.detail}",
                         request=response.request, response=response)


def parse_sse(lines: Iterable[str]) -> Iterator[Dict[str, Any]]:
    """Minimal server-sent-events reader: yields ``{"data": <decoded JSON>}`` for data fields and
    ``{"comment": text}`` for comment lines (keep-alives); every other field is skipped."""
    for raw in lines:
        line = raw.rstrip("\r\n")
        if not line:
            continue
        if line.startswith(":"):
            yield {"comment": line[1:].strip()}
            continue
        name, sep, value = line.partition(":")
        if sep and name == "data":
            yield {"data": json.loads(value.lstrip(" "))}


@dataclass
class EventSource:
    """Iterating yields the decoded ``data`` payloads ({"path", "size", "from_peer"} per received file)."""

    response: httpx.Response
    keep_alives: int = 0
    received: List[Any] = field(default_factory=list)

    def __iter__(self) -> Iterator[Any]:
        for item in parse_sse(self.response.iter_lines()):
            if "comment" in item:
                self.keep_alives += 1
                continue
            self.received.append(item["data"])
            yield item["data"]


class Session:
    """One HTTP client bound to the job's bridge socket.  Usable as a context manager."""

    def __init__(self, socket_path: str, timeout: Optional[float] = None) -> None:
        self.socket_path, self.timeout = socket_path, timeout
        self._http = httpx.Client(transport=httpx.HTTPTransport(uds=socket_path), base_url=_BASE)

    # -- lifecycle ------------------------------------------------------------------------------
    def __enter__(self) -> "Session":
        return self

    def __exit__(self, *exc_info) -> None:
        self.close()

    def close(self) -> None:
        self._http.close()

    # -- plumbing -------------------------------------------------------------------------------
    def _call(self, route: str, payload: Any = None, method: str = "POST") -> httpx.Response:
        resp = self._http.request(method, ROUTES[route], json=payload if method == "POST" else None,
                                  timeout=self.timeout)
        if resp.is_error:
            raise BridgeRequestError(resp)
        return resp

    # -- the five routes --------------------------------------------------------------------------
    def fetch(self, resource: Any) -> Any:
        """Download ``resource`` (uri / huggingface / scheduler data) into ``artifacts/``; returns
        ``[{"path", "size"}, ...]`` relative to the work dir."""
        return self._call("fetch", resource).json()

    def send_resource(self, resource: Any, path: str) -> None:
        """Push the work-dir file ``path`` to the peers of ``resource``; returns as soon as the bridge has
        accepted the request (the copy runs in the background)."""
        self._call("send", {"resource": resource, "path": path})

    def send_status(self, status: Any) -> Any:
        """Forward a progress message to the scheduler; returns its ``{"type", "counter"}`` answer."""
        return self._call("status", status).json()

    def openapi(self) -> Any:
        return self._call("openapi", method="GET").json()

    @contextmanager
    def receive(self, resource: Any, path: str) -> Iterator[EventSource]:
        """Subscribe to files pushed by the peers of ``resource``; they are stored under ``path``."""
        request = self._http.build_request("POST", ROUTES["receive"], json={"resource": resource, "path": path},
                                           headers={"Accept": "text/event-stream"}, timeout=None)
        resp = self._http.send(request, stream=True)
        try:
            if resp.is_error:
                resp.read()
                raise BridgeRequestError(resp)
            yield EventSource(resp)
        finally:
            resp.close()
