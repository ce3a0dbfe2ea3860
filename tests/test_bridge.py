"""Job-bridge route tests (the reference has none for crates/worker/src/executor/bridge.rs): the five
routes over a real Unix socket, driven by the executor-side `Session` client -- JSON shapes, 0600 files,
path rules, error statuses, the SSE receive stream and status forwarding to the scheduler."""
import asyncio
import json
import os
import stat
import threading

import pytest

from hypha_b200 import messages as M
from hypha_b200.executor.api import Session
from hypha_b200.transport.network import Network
from hypha_b200.worker.bridge import Bridge, Connector


def _run_with_bridge(tmp, scenario):
    """Start worker + scheduler + peer networks and a bridge; run the blocking ``scenario(session, ctx)``
    in a thread (the Session client is synchronous, like the reference's)."""
    async def main():
        worker, sched, peer = Network("worker"), Network("sched"), Network("peer")
        addr = await worker.listen("memory://bridge-test-worker")
        await sched.dial(addr)
        await peer.dial(addr)
        seen = []

        async def progress(p, msg):
            req = M.ProgressRequest.from_wire(msg)
            seen.append((p, req.job_id, req.progress.kind, req.progress.batch_size))
            if req.progress.kind == "status":
                return M.ProgressResponse("ScheduleUpdate", 3)
            return M.ProgressResponse("Ok", 0)

        sched.on(M.PROGRESS_PROTOCOL, None, progress)
        work = os.path.join(tmp, "work")
        import tempfile

        sock_dir = tempfile.mkdtemp(prefix="hbb", dir="/tmp")      # AF_UNIX paths are limited to 107 bytes
        sock = os.path.join(sock_dir, "b.sock")
        registry = os.path.join(tmp, "registry")
        os.makedirs(os.path.join(registry, "org", "model"))
        open(os.path.join(registry, "org", "model", "config.json"), "w").write('{"model_type": "gpt2"}')
        bridge = Bridge(Connector(worker, hf_local=registry), worker, work, sock, "job-1", "sched")
        await bridge.start()
        ctx = {"work": work, "sock": sock, "seen": seen, "peer": peer, "worker": worker, "loop": asyncio.get_event_loop()}
        try:
            with Session(sock) as session:
                await asyncio.get_event_loop().run_in_executor(None, scenario, session, ctx)
        finally:
            await bridge.stop()
            for n in (worker, sched, peer):
                await n.close()
        return ctx

    return asyncio.run(main())


def test_bridge_socket_fetch_status_and_openapi(tmp_path):
    def scenario(session, ctx):
        assert stat.S_IMODE(os.stat(ctx["sock"]).st_mode) == 0o600               # bridge.rs:154-175
        api = session.openapi()
        assert set(api["paths"]) == {"/resources/fetch", "/resources/send", "/resources/receive", "/status/send",
                                     "/openapi.json"}
        got = session.fetch(M.Fetch.huggingface("org/model", None, ["config.json"], None).to_wire())
        assert got == [{"path": "artifacts/config.json", "size": 22}]
        dst = os.path.join(ctx["work"], "artifacts", "config.json")
        assert json.load(open(dst)) == {"model_type": "gpt2"} and stat.S_IMODE(os.stat(dst).st_mode) == 0o600
        resp = session.send_status({"status": {"batch_size": 7, "round_time": -1}})     # extra key ignored (Q2)
        assert resp == {"type": "ScheduleUpdate", "counter": 3}
        assert session.send_status("update")["type"] == "Ok"
        assert session.send_status({"metrics": {"round": 2, "metrics": {"loss": 1.5}}})["type"] == "Ok"

    ctx = _run_with_bridge(str(tmp_path), scenario)
    kinds = [(p, j, k) for p, j, k, _ in ctx["seen"]]
    assert kinds == [("worker", "job-1", "status"), ("worker", "job-1", "update"), ("worker", "job-1", "metrics")]
    assert ctx["seen"][0][3] == 7
    assert not os.path.exists(ctx["sock"])                                       # removed on stop


def test_bridge_rejects_bad_requests(tmp_path):
    import httpx

    def scenario(session, ctx):
        def status_of(fn):
            with pytest.raises(httpx.HTTPStatusError) as ei:
                fn()
            return ei.value.response.status_code, ei.value.response.json()["error"]

        code, err = status_of(lambda: session.fetch({"type": "uri", "value": "ftp://example.org/x"}))
        assert code >= 400 and "http/https" in err
        code, err = status_of(lambda: session.fetch({"type": "uri", "value": "file:///etc/passwd"}))
        assert code >= 400 and "http/https" in err                     # a job spec cannot read local files
        for repo in ("../../etc", "/etc", "org/../../../etc"):         # ... nor leave the offline model registry
            code, err = status_of(lambda repo=repo: session.fetch(
                M.Fetch.huggingface(repo, None, ["passwd"], None).to_wire()))
            assert code >= 400, repo
        assert not os.path.exists(os.path.join(ctx["work"], "artifacts", "passwd"))
        code, err = status_of(lambda: session.fetch(M.Fetch.data_peer("peer", M.DataSlice("d", 0)).to_wire()))
        assert code >= 400 and "peers" in err
        code, err = status_of(lambda: session.fetch({"type": "telepathy"}))
        assert code >= 400
        os.makedirs(ctx["work"], exist_ok=True)
        open(os.path.join(ctx["work"], "ok.bin"), "wb").write(b"x")
        send = M.Send.to_peers(["peer"]).to_wire()
        for bad in ("/etc/passwd", "../outside", "a/../../b", "missing.bin"):        # bridge.rs:330-346
            code, err = status_of(lambda bad=bad: session.send_resource(send, bad))
            assert code >= 400, bad
        code, err = status_of(lambda: session.send_resource(M.Fetch.uri("http://x").to_wire(), "ok.bin"))
        assert code >= 400 and "Peers" in err
        code, err = status_of(lambda: session.send_status("upgrade"))
        assert code >= 400 and "unknown progress" in err

    _run_with_bridge(str(tmp_path), scenario)


def test_bridge_send_and_receive_streams(tmp_path):
    """send_resource returns at once and pushes in the background; receive yields one SSE event per
    incoming file from an allow-listed peer with {path, size, from_peer}; other peers are ignored."""
    def scenario(session, ctx):
        loop, peer, worker = ctx["loop"], ctx["peer"], ctx["worker"]
        inbox = asyncio.run_coroutine_threadsafe(_register(peer), loop).result(5)
        os.makedirs(ctx["work"], exist_ok=True)
        open(os.path.join(ctx["work"], "delta.bin"), "wb").write(b"d" * 3000)
        session.send_resource(M.Send.to_peers(["peer"]).to_wire(), "delta.bin")
        st = asyncio.run_coroutine_threadsafe(asyncio.wait_for(inbox.get(), 5), loop).result(6)
        data = asyncio.run_coroutine_threadsafe(_drain(st), loop).result(5)
        assert st.peer == "worker" and data == b"d" * 3000

        events = []

        def reader():
            with session.receive(M.Receive.from_peers(["peer"]).to_wire(), "incoming") as src:
                for ev in src:
                    events.append(ev)
                    return

        t = threading.Thread(target=reader)
        t.start()
        import time

        time.sleep(0.3)                                    # let /resources/receive register the receiver
        asyncio.run_coroutine_threadsafe(worker.stream_push("worker", b"ignored-self", job_id="job-1"), loop)
        asyncio.run_coroutine_threadsafe(peer.stream_push("worker", b"u" * 5000, job_id="job-1"), loop).result(5)
        t.join(10)
        assert not t.is_alive() and len(events) == 1
        ev = events[0]
        assert ev["from_peer"] == "peer" and ev["size"] == 5000 and ev["path"].startswith("incoming/")
        path = os.path.join(ctx["work"], ev["path"])
        assert open(path, "rb").read() == b"u" * 5000 and stat.S_IMODE(os.stat(path).st_mode) == 0o600

    async def _register(peer):
        return peer.streams_push("job-1")

    async def _drain(st):
        return b"".join([c async for c in st.chunks()])

    _run_with_bridge(str(tmp_path), scenario)


def test_connector_fans_out_concurrently_and_tolerates_a_dead_peer(tmp_path):
    """`Connector.send` to several peers: one stream per peer at the same time; with `tolerate_partial` an
    unreachable peer is reported in `last_send_failures` instead of failing the reachable ones; without it,
    or when nobody can be reached, the error is raised."""
    import asyncio

    from hypha_b200 import messages as M
    from hypha_b200.transport import Network, NetworkError
    from hypha_b200.worker import Connector

    async def main():
        hub = Network("hub")
        addr = await hub.listen(f"memory://fanout-{id(hub)}")
        peers = {}
        for name in ("a", "b"):
            n = Network(name)
            await n.dial(addr)
            peers[name] = (n, n.streams_push("job"))
        await asyncio.sleep(0.05)
        conn = Connector(hub)
        src = str(tmp_path / "update.bin")
        payload = os.urandom(300_000)
        open(src, "wb").write(payload)
        ref = M.Send.to_peers(["a", "ghost", "b"])
        sent = await conn.send(ref, src, job_id="job", tolerate_partial=True)
        assert sent == 2 * len(payload)
        assert [p for p, _ in conn.last_send_failures] == ["ghost"]
        for name in ("a", "b"):
            st = await asyncio.wait_for(peers[name][1].get(), 5)
            assert await st.read_all() == payload and st.peer == "hub"
        with pytest.raises(NetworkError):
            await conn.send(ref, src, job_id="job")                     # strict: any failure raises
        with pytest.raises(NetworkError):
            await conn.send(M.Send.to_peers(["ghost", "phantom"]), src, job_id="job", tolerate_partial=True)
        for n, _ in peers.values():
            await n.close()
        await hub.close()
    asyncio.run(main())
