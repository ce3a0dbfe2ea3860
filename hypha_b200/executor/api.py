"""``Session``: the executor-side client of the job bridge (HTTP over a Unix socket).

Same surface as the reference client (executors/accelerate/src/hypha/accelerate_executor/api.py:10-62):
``fetch`` / ``send_resource`` / ``send_status`` / ``receive`` (context manager yielding an SSE
``EventSource``)."""
from __future__ import annotations

import json
from contextlib import AbstractContextManager, contextmanager
from typing import Any, Iterator

import httpx


class Session(AbstractContextManager):
    def __init__(self, socket_path: str) -> None:
        self._client = httpx.Client(transport=httpx.HTTPTransport(uds=socket_path))

    def __exit__(self, exc_type, exc_value, traceback) -> None:
        self._client.close()

    def close(self) -> None:
        self._client.close()

    def send_resource(self, resource: Any, path: str) -> None:
        self._client.post("http://hypha/resources/send", json={"resource": resource, "path": path},
                          timeout=None).raise_for_status()

    def send_status(self, status: Any) -> Any:
        return self._client.post("http://hypha/status/send", json=status, timeout=None).raise_for_status().json()

    def fetch(self, resource: Any) -> Any:
        return self._client.post("http://hypha/resources/fetch", json=resource, timeout=None).raise_for_status().json()

    def openapi(self) -> Any:
        return self._client.get("http://hypha/openapi.json").raise_for_status().json()

    @contextmanager
    def receive(self, resource: Any, path: str) -> Iterator["EventSource"]:
        with self._client.stream("POST", "http://hypha/resources/receive", json={"resource": resource, "path": path},
                                 headers={"Accept": "text/event-stream"}, timeout=None) as resp:
            yield EventSource(resp)


class EventSource:
    def __init__(self, response: httpx.Response) -> None:
        self.response = response

    def __iter__(self) -> Iterator[Any]:
        for line in self.response.iter_lines():
            name, _, value = line.rstrip("\n").partition(":")
            if name == "data":
                yield json.loads(value)
            # comments (keep-alives) and other SSE fields are ignored
