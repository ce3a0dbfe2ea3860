"""Connector (fetch / send / receive routing) and the per-job HTTP-over-UDS bridge.

Reference: crates/worker/src/connector/mod.rs (:65-87 traits, :108-195 router, :226-302
HttpHfFetcher, :318-433 push connector, :443-507 pull connector) and
crates/worker/src/executor/bridge.rs (:154-160 routes, :216-248 fetch, :256-327 send,
:330-346 path safety, :350-377 validation, :392-504 SSE receive, :506-523 status).

Same five routes, same JSON shapes, same path rules (relative only, no ``..``, files 0600).
The executor process talks to it with the ``Session`` client (hypha_b200/executor/api.py), or
with the reference's own httpx client -- the wire format is identical.
"""
from __future__ import annotations

import asyncio
import hashlib
import json
import os
import random
import shutil
from typing import Any, AsyncIterator, Callable, Dict, Optional, Tuple
from urllib.parse import urlparse

from aiohttp import web

from .. import messages as M
from ..transport import IncomingStream, Network, NetworkError


class BridgeError(Exception):
    status = 500


class InvalidPath(BridgeError):
    status = 400


class InvalidResource(BridgeError):
    status = 400


def safe_join(base: str, rel: str) -> str:
    """Join ``rel`` under ``base``; absolute paths and ``..`` components are rejected."""
    if not rel or os.path.isabs(rel):
        raise InvalidPath(f"path must be relative: {rel!r}")
    parts = rel.replace("\\", "/").split("/")
    if any(p == ".." for p in parts):
        raise InvalidPath(f"path must not contain '..': {rel!r}")
    out = os.path.normpath(os.path.join(base, *parts))
    if not (out == base or out.startswith(base.rstrip("/") + "/")):
        raise InvalidPath(f"path escapes the work dir: {rel!r}")
    return out


# ---------------------------------------------------------------------------------------------
# connector
# ---------------------------------------------------------------------------------------------
class Connector:
    """Routes ``Fetch`` / ``Send`` / ``Receive`` references to transports."""

    def __init__(self, network: Network, *, hf_local: Optional[str] = None):
        self.network = network
        # offline model registry: <hf_local>/<repository>/<filename>; HYPHA_HF_LOCAL overrides
        self.hf_local = hf_local or os.environ.get("HYPHA_HF_LOCAL")
        self.last_send_failures: list = []      # [(peer, reason)] of the last tolerate_partial send

    # -- fetch: yields (file_name, async chunk iterator | local path) ---------------------------
    async def fetch(self, ref: M.Reference, scheduler_hint: Optional[str] = None):
        if ref.type == "uri":
            yield await self._fetch_uri(ref.value)
        elif ref.type == "huggingface":
            async for item in self._fetch_hf(ref):
                yield item
        elif ref.type == "scheduler":
            yield await self._fetch_slice(ref.peer, ref.dataset)
        elif ref.type == "peers":
            if ref.resource is None or not ref.peers:
                raise InvalidResource("peers fetch needs a resource and one peer")
            st = await self.network.stream_pull(ref.peers[0], ref.resource)
            yield (str(ref.resource.index), st)
        else:
            raise InvalidResource(f"cannot fetch {ref.type}")

    async def _fetch_uri(self, url: str) -> Tuple[str, Any]:
        u = urlparse(url)
        name = os.path.basename(u.path) or "download"
        if u.scheme == "file":
            return name, u.path
        if u.scheme not in ("http", "https"):
            raise InvalidResource(f"unsupported URI scheme {u.scheme!r}")
        import aiohttp

        async def chunks():
            async with aiohttp.ClientSession() as s:
                async with s.get(url) as r:
                    r.raise_for_status()
                    async for c in r.content.iter_chunked(1 << 20):
                        yield c

        return name, chunks()

    async def _fetch_hf(self, ref: M.Reference):
        for fn in ref.filenames:
            local = None
            if self.hf_local:
                # repository and file name come from the scheduler's job spec: they must stay inside the registry
                cand = safe_join(self.hf_local, f"{ref.repository}/{fn}")
                if os.path.exists(cand):
                    local = cand
            if local is None:
                try:
                    from huggingface_hub import hf_hub_download

                    local = await asyncio.get_event_loop().run_in_executor(
                        None, lambda: hf_hub_download(ref.repository, fn, revision=ref.revision, token=ref.token))
                except Exception as exc:
                    raise BridgeError(f"cannot fetch {ref.repository}/{fn}: {exc} "
                                      f"(offline? place the file under $HYPHA_HF_LOCAL/{ref.repository}/)") from exc
            yield (fn, local)

    async def _fetch_slice(self, scheduler: str, dataset: str) -> Tuple[str, Any]:
        """Ask the scheduler which slice is next, then pull it from the data node
        (connector/mod.rs:457-506)."""
        resp = M.ApiResponse.from_wire(
            await self.network.request(scheduler, M.API_PROTOCOL, M.ApiRequest.data(dataset)))
        kind, body = resp.variant()
        if resp.kind != "Data" or kind != "Success":
            raise BridgeError(f"scheduler has no data for {dataset!r}: {kind}")
        provider, index = body["data_provider"], int(body["index"])
        st = await self.network.stream_pull(provider, M.DataSlice(dataset, index))
        return str(index), st

    # -- send -----------------------------------------------------------------------------------
    async def send(self, ref: M.Reference, path: str, job_id: Optional[str] = None, tolerate_partial: bool = False) -> int:
        """Push ``path`` to the reference's peers; returns the bytes sent.  With ``tolerate_partial`` a peer that
        cannot be reached does not fail the others (``last_send_failures`` lists them); it still raises when
        nobody could be reached."""
        if ref.type != "peers":
            raise InvalidResource("send needs a peers reference")
        peers = list(ref.peers)
        strat = ref.strategy or M.SelectionStrategy.ALL
        if strat == M.SelectionStrategy.ONE:
            peers = peers[:1]
        elif strat == M.SelectionStrategy.RANDOM:
            peers = [random.choice(peers)] if peers else []
        # one stream per peer, all at once: the links are independent sockets, so the fan-out of an update to N
        # workers takes about as long as the slowest transfer instead of the sum
        results = await asyncio.gather(*(self.network.stream_push(p, path, job_id=job_id) for p in peers),
                                       return_exceptions=True)
        failed = [(p, r) for p, r in zip(peers, results) if isinstance(r, BaseException)]
        for p, r in failed:
            if not isinstance(r, NetworkError):
                raise r
        if failed and (not tolerate_partial or len(failed) == len(peers)):
            raise failed[0][1]
        self.last_send_failures = [(p, str(r)) for p, r in failed]
        return sum(r for r in results if isinstance(r, int))

    # -- receive ------------------------------------------------------------------------------------
    async def receive(self, ref: M.Reference, job_id: Optional[str] = None,
                      allowed: Optional[set] = None) -> AsyncIterator[IncomingStream]:
        """Incoming push streams from the allow-listed peers (connector/mod.rs:397-432).  ``allowed`` may
        be a live set the caller extends when the job's membership grows."""
        if ref.type != "peers":
            raise InvalidResource("receive needs a peers reference")
        if allowed is None:
            allowed = set(ref.peers)
        q = self.network.streams_push(job_id)
        try:
            while True:
                st: IncomingStream = await q.get()
                if st.peer not in allowed:
                    st.discard()      # not on the allow-list: its sender runs out of flow-control credit and stops
                    continue
                yield st
        finally:
            self.network.close_streams_push(job_id)


# ---------------------------------------------------------------------------------------------
# bridge
# ---------------------------------------------------------------------------------------------
OPENAPI = {
    "openapi": "3.0.0", "info": {"title": "hypha job bridge", "version": "0.0.1"},
    "paths": {"/resources/fetch": {"post": {}}, "/resources/send": {"post": {}},
              "/resources/receive": {"post": {}}, "/status/send": {"post": {}}, "/openapi.json": {"get": {}}},
}


class Bridge:
    """aiohttp app on ``<work_dir>/bridge.sock`` (0600) serving one job."""

    def __init__(self, connector: Connector, network: Network, work_dir: str, socket_path: str,
                 job_id: str, scheduler: str, log: Optional[Callable[[str], None]] = None):
        self.connector, self.network = connector, network
        self.log = log or (lambda m: None)
        self.work_dir = os.path.abspath(work_dir)
        self.socket_path, self.job_id, self.scheduler = socket_path, job_id, scheduler
        self._runner: Optional[web.AppRunner] = None
        self._bg: set = set()
        self.app = web.Application(client_max_size=16 << 20)
        self.app.add_routes([
            web.post("/resources/fetch", self._fetch), web.post("/resources/send", self._send),
            web.post("/resources/receive", self._receive), web.post("/status/send", self._status),
            web.get("/openapi.json", self._openapi),
        ])

    async def start(self) -> None:
        os.makedirs(self.work_dir, exist_ok=True)
        if os.path.exists(self.socket_path):
            os.unlink(self.socket_path)
        self._runner = web.AppRunner(self.app, handle_signals=False)
        await self._runner.setup()
        site = web.UnixSite(self._runner, self.socket_path)
        await site.start()
        os.chmod(self.socket_path, 0o600)

    async def stop(self) -> None:
        for t in list(self._bg):
            t.cancel()
        if self._runner is not None:
            await self._runner.cleanup()
            self._runner = None
        try:
            os.unlink(self.socket_path)
        except OSError:
            pass

    def _spawn(self, coro) -> None:
        t = asyncio.ensure_future(coro)
        self._bg.add(t)
        t.add_done_callback(self._bg.discard)

    @staticmethod
    def _fail(exc: Exception) -> web.Response:
        status = getattr(exc, "status", 500)
        return web.json_response({"error": f"{type(exc).__name__}: {exc}"}, status=status)

    async def _store(self, name: str, src: Any, dir_rel: str) -> Dict[str, Any]:
        rel = f"{dir_rel}/{name}"
        dst = safe_join(self.work_dir, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if isinstance(src, str):
            shutil.copyfile(src, dst)
            os.chmod(dst, 0o600)
            size = os.path.getsize(dst)
        elif isinstance(src, IncomingStream):
            size = await src.save(dst)
        else:
            fd = os.open(dst, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o600)
            size = 0
            with os.fdopen(fd, "wb") as f:
                async for c in src:
                    f.write(c)
                    size += len(c)
        return {"path": rel, "size": size}

    # -- routes -----------------------------------------------------------------------------------
    async def _fetch(self, request: web.Request) -> web.Response:
        try:
            ref = M.Fetch.from_wire(await request.json())
            scheme = urlparse(ref.value or "").scheme if ref.type == "uri" else ""
            if ref.type == "uri" and scheme not in ("http", "https") and not (
                    scheme == "file" and os.environ.get("HYPHA_ALLOW_FILE_URI") == "1"):
                # bridge.rs:350-377.  file:// would let a remote scheduler's job spec copy any local file into the
                # job directory; it exists for offline tests only, behind an explicit worker-side switch
                raise InvalidResource("uri fetch supports http/https")
            if ref.type == "peers":
                raise InvalidResource("peers references cannot be fetched through the bridge")
            out = []
            async for name, src in self.connector.fetch(ref):
                out.append(await self._store(name, src, "artifacts"))
            return web.json_response(out)
        except (BridgeError, M.MessageError, NetworkError, KeyError, ValueError) as exc:
            return self._fail(exc)

    async def _send(self, request: web.Request) -> web.Response:
        try:
            body = await request.json()
            ref = M.Send.from_wire(body["resource"])
            path = safe_join(self.work_dir, body["path"])
            if not os.path.isfile(path):
                raise InvalidPath(f"no such file: {body['path']!r}")
        except (BridgeError, M.MessageError, KeyError, ValueError) as exc:
            return self._fail(exc)

        async def bg():
            try:
                await self.connector.send(ref, path, job_id=self.job_id)
            except Exception as exc:            # the HTTP request already returned 200: report it in the log
                self.log(f"[bridge {self.job_id[:8]}] sending {body['path']!r} failed: {type(exc).__name__}: {exc}")

        self._spawn(bg())                     # 200 immediately, copy runs in the background
        return web.Response(status=200)

    async def _receive(self, request: web.Request) -> web.StreamResponse:
        try:
            body = await request.json()
            ref = M.Receive.from_wire(body["resource"])
            dir_rel = body.get("path") or "incoming"
            os.makedirs(safe_join(self.work_dir, dir_rel), exist_ok=True)
        except (BridgeError, M.MessageError, KeyError, ValueError) as exc:
            return self._fail(exc)
        resp = web.StreamResponse(status=200, headers={"Content-Type": "text/event-stream",
                                                        "Cache-Control": "no-cache"})
        await resp.prepare(request)
        events: asyncio.Queue = asyncio.Queue(64)

        async def pump():
            idx = 0
            async for st in self.connector.receive(ref, job_id=self.job_id):
                name = hashlib.sha256(f"{st.peer}:{idx}".encode()).hexdigest()[:32]
                try:
                    info = await self._store(name, st, dir_rel)
                except NetworkError as exc:       # the sender vanished mid-file: never announce half a tensor file
                    self.log(f"[bridge] incoming file from {st.peer} dropped: {exc}")
                    partial = safe_join(self.work_dir, f"{dir_rel}/{name}")
                    if os.path.exists(partial):
                        os.unlink(partial)
                    continue
                info["from_peer"] = st.peer
                await events.put(info)
                idx += 1

        task = asyncio.ensure_future(pump())
        self._bg.add(task)
        try:
            while True:
                try:
                    ev = await asyncio.wait_for(events.get(), timeout=5.0)
                    await resp.write(f"data:{json.dumps(ev)}\n\n".encode())
                except asyncio.TimeoutError:
                    await resp.write(b": keep-alive\n\n")       # every 5 s (bridge.rs:392-504)
        except (ConnectionResetError, asyncio.CancelledError, ConnectionError):
            pass
        finally:
            task.cancel()
            self._bg.discard(task)
        return resp

    async def _status(self, request: web.Request) -> web.Response:
        try:
            progress = M.Progress.from_wire(await request.json())
            # `update` of a collective job is answered when every member reached its sync point: no short
            # RPC timeout on it (a dropped rank releases the barrier through the scheduler, not a timer)
            timeout = 6 * 3600.0 if progress.kind == "update" else 30.0
            out = await self.network.request(self.scheduler, M.PROGRESS_PROTOCOL,
                                             M.ProgressRequest(self.job_id, progress), timeout=timeout)
            return web.json_response(M.ProgressResponse.from_wire(out).to_wire())
        except (M.MessageError, NetworkError, KeyError, ValueError) as exc:
            return self._fail(exc)

    async def _openapi(self, request: web.Request) -> web.Response:
        return web.json_response(OPENAPI)
