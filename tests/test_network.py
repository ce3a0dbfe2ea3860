"""Multi-node-without-a-cluster tests of the control-plane fabric (model: the reference's
crates/network/tests/{gossipsub,kad,request_response}_test.rs, 2-3 in-memory nodes per test)."""
import asyncio
import os
import tempfile

import pytest

from hypha_b200.transport import (AlreadyRegistered, DialError, Network, NoHandler, PublishError,
                                  RecordNotFound, RequestFailed)


def run(coro):
    return asyncio.run(asyncio.wait_for(coro, 20))


async def pair(kind="memory", tmp=None):
    a, b = Network("A"), Network("B")
    if kind == "memory":
        addr = await a.listen(f"memory://{id(a)}")
    elif kind == "unix":
        addr = await a.listen(f"unix://{tmp}/a.sock")
    else:
        addr = await a.listen("tcp://127.0.0.1:0")
    assert await b.dial(addr) == "A"
    for _ in range(50):
        if "B" in a.links:
            break
        await asyncio.sleep(0.01)
    return a, b


@pytest.mark.parametrize("kind", ["memory", "unix", "tcp"])
def test_request_response_roundtrip(kind):
    async def main():
        with tempfile.TemporaryDirectory() as tmp:
            a, b = await pair(kind, tmp)

            async def h(peer, msg):
                return {"echo": msg["x"], "from": peer}

            a.on("/hypha-api/0.0.1", None, h)
            r = await b.request("A", "/hypha-api/0.0.1", {"x": 7})
            assert r == {"echo": 7, "from": "B"}
            assert b.counters.total()["hypha.bandwidth.outbound.bytes"] > 0
            assert a.counters.snapshot()["hypha.bandwidth.inbound.bytes"].get("/hypha-api/0.0.1", 0) > 0
            await a.close(); await b.close()
    run(main())


def test_multiple_handlers_with_patterns_first_match_wins():
    async def main():
        a, b = await pair()

        async def h1(p, m):
            return "one"

        async def h2(p, m):
            return "two"

        async def h3(p, m):
            return "any"

        a.on("/p", lambda m: m.get("k") == 1, h1)
        a.on("/p", lambda m: m.get("k") == 2, h2)
        a.on("/p", lambda m: True, h3)
        a.on("/p", lambda m: m.get("k") == 2, h1)         # duplicate pattern registered later: never wins
        assert await b.request("A", "/p", {"k": 1}) == "one"
        assert await b.request("A", "/p", {"k": 2}) == "two"
        assert await b.request("A", "/p", {"k": 3}) == "any"
        await a.close(); await b.close()
    run(main())


def test_handler_unregistration_and_no_handler():
    async def main():
        a, b = await pair()

        async def h(p, m):
            return "ok"

        reg = a.on("/p", None, h)
        assert await b.request("A", "/p", {}) == "ok"
        reg.unregister()
        with pytest.raises(NoHandler):
            await b.request("A", "/p", {})
        with pytest.raises(NoHandler):
            await b.request("A", "/other", {})
        await a.close(); await b.close()
    run(main())


def test_handler_exception_is_failed_request():
    async def main():
        a, b = await pair()

        async def h(p, m):
            raise ValueError("boom")

        a.on("/p", None, h)
        with pytest.raises(RequestFailed):
            await b.request("A", "/p", {})
        await a.close(); await b.close()
    run(main())


def test_request_after_disconnect():
    """request_response_test.rs:744-792: peer goes away -> Err(Request)."""
    async def main():
        a, b = await pair()

        async def h(p, m):
            return 1

        a.on("/p", None, h)
        assert await b.request("A", "/p", {}) == 1
        await a.close()
        with pytest.raises(RequestFailed):
            await b.request("A", "/p", {}, timeout=1.0)
        await b.close()
    run(main())


def test_dial_errors():
    async def main():
        n = Network("X")
        with pytest.raises(DialError):
            await n.dial("memory://nobody")
        with pytest.raises(DialError):
            await n.dial("tcp://127.0.0.1:1", timeout=1.0)
        await n.close()
    run(main())


def test_gossip_single_and_multiple_subscribers():
    async def main():
        hub = Network("hub", relay_server=True, dht_server=True)
        addr = await hub.listen(f"memory://hub-{id(hub)}")
        pub, s1, s2 = Network("pub"), Network("s1"), Network("s2")
        for n in (pub, s1, s2):
            await n.dial(addr)
        sub1 = await s1.subscribe("hypha/worker")
        sub2 = await s2.subscribe("hypha/worker")
        await pub.publish("hypha/worker", b"hello")
        assert await sub1.get(2) == ("pub", b"hello")
        assert await sub2.get(2) == ("pub", b"hello")
        # publisher does not hear itself; unsubscribed peers stop receiving
        await sub2.unsubscribe()
        await pub.publish("hypha/worker", b"again")
        assert (await sub1.get(2))[1] == b"again"
        assert sub2.q.empty()
        for n in (hub, pub, s1, s2):
            await n.close()
    run(main())


def test_isolated_peer_publish_fails():
    async def main():
        n = Network("lonely")
        with pytest.raises(PublishError):
            await n.publish("hypha/worker", b"x")
        await n.close()
    run(main())


def test_store_and_get_record_and_providers():
    async def main():
        gw = Network("gw", relay_server=True, dht_server=True)
        gw.enable_dht_handlers()
        addr = await gw.listen(f"memory://gw-{id(gw)}")
        data, sched = Network("data"), Network("sched")
        await data.dial(addr)
        await sched.dial(addr)
        await sched.wait_for_bootstrap(1.0)
        await data.put_record("mnist", b'{"num_slices": 7}')
        val, publisher = await sched.get_record("mnist")
        assert val == b'{"num_slices": 7}' and publisher == "data"
        with pytest.raises(RecordNotFound):
            await sched.get_record("missing")
        await data.provide("mnist")
        lonely = Network("p2")
        await lonely.dial(addr)
        await lonely.provide("mnist")
        assert await sched.find_providers("mnist") == ["data", "p2"]
        assert set(gw.closest_peers("mnist")) == {"data", "sched", "p2"}
        for n in (gw, data, sched, lonely):
            await n.close()
    run(main())


def test_wait_for_bootstrap_times_out_without_gateway():
    async def main():
        n = Network("n")
        with pytest.raises(asyncio.TimeoutError):
            await n.wait_for_bootstrap(0.1)
        await n.close()
    run(main())


def test_relay_through_gateway():
    """Two peers that only know the gateway can still talk (relay server)."""
    async def main():
        gw = Network("gw", relay_server=True, dht_server=True)
        addr = await gw.listen(f"memory://gw-{id(gw)}")
        a, b = Network("a"), Network("b")
        await a.dial(addr)
        await b.dial(addr)

        async def h(p, m):
            return {"hi": p}

        b.on("/p", None, h)
        assert await a.request("b", "/p", {}) == {"hi": "a"}
        await a.ping("b")
        for n in (gw, a, b):
            await n.close()
    run(main())


def test_direct_dial_via_gateway_peer_table():
    async def main():
        gw = Network("gw", relay_server=True, dht_server=True)
        gaddr = await gw.listen(f"memory://gw-{id(gw)}")
        b = Network("b")
        await b.listen(f"memory://b-{id(b)}")
        await b.dial(gaddr)
        a = Network("a")
        await a.dial(gaddr)          # learns b's address from the gateway's hello

        async def h(p, m):
            return "direct"

        b.on("/p", None, h)
        assert await a.request("b", "/p", {}) == "direct"
        assert "b" in a.links        # a direct link was established
        for n in (gw, a, b):
            await n.close()
    run(main())


@pytest.mark.parametrize("kind", ["memory", "tcp"])
def test_push_stream_and_already_registered(kind):
    async def main():
        with tempfile.TemporaryDirectory() as tmp:
            a, b = await pair(kind, tmp)
            rx = a.streams_push()
            with pytest.raises(AlreadyRegistered):
                a.streams_push()
            rx_job = a.streams_push(job_id="job-1")          # per-job receivers may coexist (Q15)
            payload = os.urandom(3 * (1 << 20) + 17)
            src = os.path.join(tmp, "src.bin")
            open(src, "wb").write(payload)
            t = asyncio.ensure_future(b.stream_push("A", src))
            st = await rx.get()
            dst = os.path.join(tmp, "dst.bin")
            n = await st.save(dst)
            assert await t == len(payload) == n and open(dst, "rb").read() == payload
            assert oct(os.stat(dst).st_mode & 0o777) == "0o600" and st.peer == "B"
            t = asyncio.ensure_future(b.stream_push("A", b"tiny", job_id="job-1"))
            st2 = await rx_job.get()
            assert await st2.read_all() == b"tiny" and st2.job_id == "job-1"
            await t
            await a.close(); await b.close()
    run(main())


def test_pull_stream():
    async def main():
        a, b = await pair()
        blobs = {0: b"slice-zero" * 1000, 1: b"slice-one"}

        async def serve(peer, header):
            return blobs[header["index"]]

        a.serve_pull(serve)
        with pytest.raises(AlreadyRegistered):
            a.serve_pull(serve)
        st = await b.stream_pull("A", {"dataset": "d", "index": 0})
        assert await st.read_all() == blobs[0]
        st = await b.stream_pull("A", {"dataset": "d", "index": 1})
        assert await st.read_all() == blobs[1]
        await a.close(); await b.close()
    run(main())


def test_pull_streams_in_service_are_limited_per_peer():
    """A data node serves at most `STREAM_ACCEPT_LIMIT` pull streams of one peer at a time (stream_pull.rs:71); the
    next one is refused with a reason, and the slot is free again when one of them finishes."""
    from hypha_b200.transport.network import STREAM_ACCEPT_LIMIT, StreamAborted

    async def main():
        a, b = await pair()
        gate = asyncio.Event()

        async def serve(peer, header):
            await gate.wait()
            return b"x" * 10

        a.serve_pull(serve)
        held = [await b.stream_pull("A", {"i": i}) for i in range(STREAM_ACCEPT_LIMIT)]
        await asyncio.sleep(0.05)
        assert a._pulls_active == {"B": STREAM_ACCEPT_LIMIT}
        extra = await b.stream_pull("A", {"i": "one too many"})
        with pytest.raises(StreamAborted, match="pull streams at once"):
            await extra.read_all()
        gate.set()
        for st in held:
            assert await st.read_all() == b"x" * 10
        await asyncio.sleep(0.05)
        assert a._pulls_active == {}
        assert await (await b.stream_pull("A", {"i": "again"})).read_all() == b"x" * 10
        await a.close(); await b.close()
    run(main())


def test_batched_count_and_time_windows():
    from hypha_b200.transport.utils import Batched

    async def main():
        q: asyncio.Queue = asyncio.Queue()

        async def src():
            while True:
                x = await q.get()
                if x is None:
                    return
                yield x

        out = []

        async def consume():
            async for b in Batched(src(), 3, 0.1):
                out.append(b)

        t = asyncio.ensure_future(consume())
        for i in range(4):
            q.put_nowait(i)                      # 3 -> full batch, 1 left pending
        await asyncio.sleep(0.25)                # window elapses -> partial batch flushed
        q.put_nowait(9)
        q.put_nowait(None)
        await t
        assert out == [[0, 1, 2], [3], [9]]
    run(main())


def test_cidr_helpers_and_identity(tmp_path):
    from hypha_b200 import certutil as cu
    from hypha_b200.transport.utils import RESERVED_IPV4, find_containing_cidr, is_excluded, peer_id_from_key_pem

    assert find_containing_cidr("10.1.2.3", RESERVED_IPV4) == "10.0.0.0/8"
    assert find_containing_cidr("8.8.8.8", RESERVED_IPV4) is None and not is_excluded("not-an-ip", RESERVED_IPV4)
    assert is_excluded("::1", ["::1/128"])
    d = str(tmp_path)
    cu.make_root("o", out_dir=d)
    cu.make_org(f"{d}/o-root-cert.pem", f"{d}/o-root-key.pem", "o", out_dir=d)
    cu.make_node(f"{d}/o-ca-cert.pem", f"{d}/o-ca-key.pem", "n1", out_dir=d)
    a, b = peer_id_from_key_pem(f"{d}/n1-key.pem"), peer_id_from_key_pem(f"{d}/n1-key.pem")
    assert a == b and a.startswith("12D3Koo") and a != peer_id_from_key_pem(f"{d}/o-ca-key.pem")


def test_tcp_links_use_mutual_tls(tmp_path):
    """tcp:// links run TLS 1.3 with client-certificate verification when the PKI files are
    configured; a peer without a certificate from the trusted chain cannot connect."""
    from hypha_b200 import certutil
    from hypha_b200.transport.network import DialError, Network

    d = str(tmp_path)
    rc, rk = certutil.make_root("Acme", out_dir=d)
    oc, ok = certutil.make_org(rc, rk, "Acme", name="acme-org", out_dir=d)
    paths = {}
    for n in ("alpha", "beta"):
        c, k, t = certutil.make_node(oc, ok, n, sans=("127.0.0.1",), out_dir=d)
        paths[n] = {"cert_pem": c, "key_pem": k, "trust_pem": t}
    # an unrelated PKI
    d2 = str(tmp_path / "evil")
    import os
    os.makedirs(d2)
    rc2, rk2 = certutil.make_root("Evil", out_dir=d2)
    oc2, ok2 = certutil.make_org(rc2, rk2, "Evil", name="evil-org", out_dir=d2)
    c, k, t = certutil.make_node(oc2, ok2, "mallory", out_dir=d2)
    evil = {"cert_pem": c, "key_pem": k, "trust_pem": t}

    from hypha_b200.transport.utils import peer_id_from_key_pem as pid

    async def main():
        # a node's peer id IS its key (roles.start_node derives it the same way)
        ida, idb = pid(paths["alpha"]["key_pem"]), pid(paths["beta"]["key_pem"])
        a = Network(ida, tls=paths["alpha"])
        b = Network(idb, tls=paths["beta"])
        m = Network(pid(evil["key_pem"]), tls=evil)
        plain = Network("plain")
        addr = await a.listen("tcp://127.0.0.1:0")

        async def echo(peer, msg):
            return {"echo": msg, "from": peer}

        a.on("/test/echo", None, echo)
        peer = await b.dial(addr)
        assert peer == ida
        resp = await b.request(ida, "/test/echo", {"x": 1})
        assert resp == {"echo": {"x": 1}, "from": idb}
        for bad in (m, plain):
            with pytest.raises(DialError):
                await bad.dial(addr, timeout=2.0)
        for n in (a, b, m, plain):
            await n.close()

    asyncio.run(main())


def test_only_a_dialled_gateway_can_relay_in_other_peers_names():
    """A relay vouches for the `src` of what it forwards, so a peer does not get that role by saying "relay": it must
    be a peer the application dialled (a configured gateway).  An inbound peer that claims to be a relay -- and a
    peer this node dialled on its own at an address it learnt from others -- cannot inject frames in somebody
    else's name; nor is a stranger's peer table (`known`) taken in."""
    import cbor2

    async def main():
        victim = Network("victim")
        addr = await victim.listen("memory://victim")
        got = []

        async def whoami(peer, msg):
            got.append(peer)
            return {"peer": peer}

        victim.on("/test/whoami", None, whoami)
        rogue = Network("rogue", relay_server=True)              # announces "relay": true in its hello
        rogue.peer_addrs["ghost"] = ["tcp://203.0.113.9:1"]       # ... and a peer table
        raddr = await rogue.listen("memory://rogue")
        await rogue.dial(addr)
        assert not victim.links["rogue"].is_relay and "ghost" not in victim.peer_addrs
        forged = {"t": "relay", "dst": "victim", "src": "scheduler",
                  "inner": {"t": "req", "id": "r1", "proto": "/test/whoami", "payload": cbor2.dumps({})}}
        before = victim.identity_rejections
        await rogue.links["victim"].send(dict(forged))
        await asyncio.sleep(0.1)
        assert got == [] and victim.identity_rejections == before + 1
        # the node reaches the rogue through an address it learnt (routing), not through configuration
        await victim.disconnect("rogue")
        await asyncio.sleep(0.05)
        victim.peer_addrs["rogue"] = [raddr]
        link, relayed = await victim._link_to("rogue")
        assert not relayed and not link.is_relay
        await rogue.links["victim"].send(dict(forged))
        await asyncio.sleep(0.1)
        assert got == [] and victim.identity_rejections == before + 2
        # a gateway the application dials keeps the role
        gw = Network("gw", relay_server=True)
        gaddr = await gw.listen("memory://gw")
        await victim.dial(gaddr)
        other = Network("other")
        await other.dial(gaddr)
        assert victim.links["gw"].is_relay
        assert (await other.request("victim", "/test/whoami", {}))["peer"] == "other" and got == ["other"]
        for n in (victim, rogue, gw, other):
            await n.close()

    asyncio.run(main())


def test_peer_identity_is_bound_to_the_certificate(tmp_path):
    """A node with a VALID certificate of the trusted CA cannot speak under another node's peer id: its hello
    is refused (the id must be the one derived from its certificate key), a relay stamps the id it
    authenticated on relayed frames, and DHT records are only accepted under the sender's own id."""
    from hypha_b200 import certutil
    from hypha_b200.transport.network import DialError, Network
    from hypha_b200.transport.utils import peer_id_from_key_pem as pid

    d = str(tmp_path)
    rc, rk = certutil.make_root("Acme", out_dir=d)
    oc, ok = certutil.make_org(rc, rk, "Acme", name="acme-org", out_dir=d)
    tls = {}
    for n in ("gw", "sched", "worker", "mallory"):
        c, k, t = certutil.make_node(oc, ok, n, sans=("127.0.0.1",), out_dir=d)
        tls[n] = {"cert_pem": c, "key_pem": k, "trust_pem": t}
    ids = {n: pid(v["key_pem"]) for n, v in tls.items()}

    async def main():
        gw = Network(ids["gw"], tls=tls["gw"], relay_server=True, dht_server=True)
        gw.enable_dht_handlers()
        sched = Network(ids["sched"], tls=tls["sched"])
        addr = await gw.listen("tcp://127.0.0.1:0")
        assert await sched.dial(addr) == ids["gw"]
        # 1. mallory holds a certificate of the same CA but claims the scheduler's id
        imp = Network(ids["sched"], tls=tls["mallory"])
        with pytest.raises(DialError):
            await imp.dial(addr, timeout=2.0)
        assert gw.identity_rejections == 1 and gw.links[ids["sched"]] is not None
        assert await sched.ping(ids["gw"]) is not None      # the honest link is unaffected
        # 2. under her own id she connects, but cannot publish a record or relay a frame in somebody else's name
        mal = Network(ids["mallory"], tls=tls["mallory"])
        assert await mal.dial(addr) == ids["gw"]
        link = mal.links[ids["gw"]]
        await link.send({"t": "dht_put", "key": "dataset", "value": b"x", "publisher": ids["sched"]})
        got = []

        async def whoami(peer, msg):
            got.append(peer)
            return {"peer": peer}

        sched.on("/test/whoami", None, whoami)
        await link.send({"t": "relay", "dst": ids["sched"], "src": ids["worker"],
                         "inner": {"t": "req", "id": "r1", "proto": "/test/whoami", "payload": __import__("cbor2").dumps({})}})
        await asyncio.sleep(0.3)
        assert "dataset" not in gw._records and gw.identity_rejections == 2
        # under a PKI the receiver believes the SENDER's signature, not the relay: an unsigned envelope is dropped
        # (the relay had stamped mallory's id on it) ...
        assert got == [] and sched.identity_rejections == 1
        # ... even a relay that plays along cannot pass her envelope off as the worker's: her signature does not
        # cover that name ...
        forged = mal._seal_relayed(ids["sched"], {"t": "req", "id": "r2", "proto": "/test/whoami",
                                                  "payload": __import__("cbor2").dumps({})})
        await gw.links[ids["sched"]].send(dict(forged, src=ids["worker"]))
        await asyncio.sleep(0.2)
        assert got == [] and sched.identity_rejections == 2
        # ... her own request through the relay is served under her own id, once: a replayed copy is dropped
        assert (await mal.request(ids["sched"], "/test/whoami", {}))["peer"] == ids["mallory"]
        sealed = mal._seal_relayed(ids["sched"], {"t": "req", "id": "r3", "proto": "/test/whoami",
                                                  "payload": __import__("cbor2").dumps({})})
        await link.send(dict(sealed))
        await link.send(dict(sealed))
        await asyncio.sleep(0.3)
        assert got == [ids["mallory"]] * 2 and sched.identity_rejections == 3
        for n in (gw, sched, imp, mal):
            await n.close()

    asyncio.run(main())


def test_nvls_descriptor_exchange_helpers(tmp_path):
    """The VMM handle exchange of parallel/nvls.py passes POSIX file descriptors between rank processes
    over AF_UNIX (SCM_RIGHTS); here with pipe fds instead of CUDA allocation handles."""
    import os
    import socket
    import threading

    from hypha_b200.parallel.nvls import _recv_fds, _send_fds

    import tempfile

    path = os.path.join(tempfile.mkdtemp(prefix="hbn", dir="/tmp"), "r0.sock")   # AF_UNIX paths: 107 bytes at most
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    srv.bind(path)
    srv.listen(4)
    srv.settimeout(10)
    box = {}
    t = threading.Thread(target=lambda: box.update(_recv_fds(srv, 2, 2)))
    t.start()
    pipes = {1: os.pipe(), 2: os.pipe()}
    for rank, (r, w) in pipes.items():
        _send_fds(path, rank, [w] if rank == 2 else [w, r])
    t.join()
    srv.close()
    assert sorted(box) == [1, 2] and len(box[1]) == 2 and len(box[2]) == 1
    os.write(box[2][0], b"hello")                      # the duplicated descriptor reaches the same pipe
    assert os.read(pipes[2][0], 5) == b"hello"
    for fds in box.values():
        for fd in fds:
            os.close(fd)
    for r, w in pipes.values():
        os.close(r)
        os.close(w)


def test_push_stream_waits_for_a_late_receiver():
    """A push stream that arrives before the job's receiver registered is parked and delivered once the
    receiver exists (a starting executor must not lose the aggregate worker's first file); streams
    without a job id are still dropped when nobody listens."""
    from hypha_b200.transport.network import Network

    async def main():
        a, b = Network("a"), Network("b")
        addr = await b.listen("memory://late-rx")
        await a.dial(addr)
        await a.stream_push("b", b"payload-1", job_id="job-7")       # nobody listens for job-7 yet
        await a.stream_push("b", b"orphan")                          # no job id: dropped
        await asyncio.sleep(0.05)
        q = b.streams_push("job-7")
        st = await asyncio.wait_for(q.get(), 2.0)
        data = b"".join([c async for c in st.chunks()])
        assert st.peer == "a" and data == b"payload-1"
        assert q.empty()
        await a.close()
        await b.close()

    asyncio.run(main())


def test_bulk_frames_keep_payload_and_count_wire_bytes():
    """Stream chunks >= 4 KiB travel as split frames (CBOR header + raw payload, no re-encoding on either
    side); smaller ones as plain CBOR.  Both arrive intact, also interleaved with requests on the same link,
    and the bandwidth counters see the bytes that were on the wire."""
    from hypha_b200.transport.network import SPLIT_FRAME_MIN

    async def main():
        with tempfile.TemporaryDirectory() as tmp:
            a, b = await pair("unix", tmp)
            a.on("/echo/1", None, lambda peer, m: _echo(m))
            rx = a.streams_push(job_id="j")
            sizes = [0, 1, SPLIT_FRAME_MIN - 1, SPLIT_FRAME_MIN, (1 << 20) - 1, (1 << 20) + 1, 5 * (1 << 20) + 3]
            for n in sizes:
                payload = os.urandom(n)
                push = asyncio.ensure_future(b.stream_push("A", payload, job_id="j"))
                echo = asyncio.ensure_future(b.request("A", "/echo/1", {"n": n}))
                st = await rx.get()
                got = await st.read_all()
                assert await push == n and got == payload, n
                assert (await echo) == {"n": n}
            snap = a.counters.snapshot()["hypha.bandwidth.inbound.bytes"]["/hypha-tensor-stream/push"]
            assert sum(sizes) < snap < sum(sizes) + 4096 * len(sizes)
            await a.close(); await b.close()

    async def _echo(m):
        return m
    run(main())


def test_garbage_on_a_socket_drops_only_that_link():
    """A peer that sends an oversized length prefix or bytes that are not CBOR is disconnected; the node keeps
    serving its other links."""
    import struct

    async def main():
        with tempfile.TemporaryDirectory() as tmp:
            a, b = await pair("unix", tmp)
            a.on("/echo/1", None, lambda peer, m: _ret(m))
            for junk in (struct.pack("<I", (1 << 30)) + b"x" * 16, struct.pack("<I", 5) + b"\xff\xff\xff\xff\xff",
                         struct.pack("<II", (1 << 31) | 6, 99) + b"zz"):
                r, w = await asyncio.open_unix_connection(f"{tmp}/a.sock")
                w.write(junk)
                await w.drain()
                try:
                    assert await asyncio.wait_for(r.read(1), 5) == b""    # closed by the node ...
                except ConnectionResetError:
                    pass                                                  # ... or reset (unread bytes pending)
                w.close()
            assert await b.request("A", "/echo/1", {"still": "here"}) == {"still": "here"}
            await a.close(); await b.close()

    async def _ret(m):
        return m
    run(main())


def test_stream_that_loses_its_sender_is_not_mistaken_for_a_complete_file():
    """The sender's link drops mid-push: the receiver's `save` raises StreamAborted instead of returning a
    truncated file; a pull whose handler fails raises on the requesting side as well."""
    from hypha_b200.transport import StreamAborted

    async def main():
        with tempfile.TemporaryDirectory() as tmp:
            a, b = await pair("unix", tmp)
            rx = a.streams_push(job_id="j")

            # a sender that opens a stream, sends one chunk and then loses its connection
            await b._send_to("A", {"t": "stream_open", "sid": "s1", "proto": "/hypha-tensor-stream/push", "job": "j",
                                   "header": None}, "/hypha-tensor-stream/push")
            await b._send_to("A", {"t": "stream_data", "sid": "s1", "data": b"x" * 100_000,
                                   "proto": "/hypha-tensor-stream/push"}, "/hypha-tensor-stream/push")
            st = await rx.get()
            saver = asyncio.ensure_future(st.save(os.path.join(tmp, "partial.bin")))
            await asyncio.sleep(0.1)
            await b.disconnect("A")
            with pytest.raises(StreamAborted, match="lost mid-stream"):
                await asyncio.wait_for(saver, 5)

            # pull: the serving side fails -> the requester sees the reason
            async def broken(peer, header):
                raise FileNotFoundError("slice 7 does not exist")
            a.serve_pull(broken)
            c = Network("C")
            await c.dial(a.listen_addrs[0])
            pulled = await c.stream_pull("A", {"dataset": "d", "index": 7})
            with pytest.raises(StreamAborted, match="slice 7 does not exist"):
                await asyncio.wait_for(pulled.read_all(), 5)
            await a.close(); await b.close(); await c.close()
    run(main())


def test_revoked_node_certificate_is_refused(tmp_path):
    """`crls_pem`: a node whose certificate is on the organisation CA's revocation list cannot connect, its
    sibling from the same CA still can."""
    from hypha_b200 import certutil
    from hypha_b200.transport.network import DialError, Network

    d = str(tmp_path)
    rc, rk = certutil.make_root("Acme", out_dir=d)
    oc, ok = certutil.make_org(rc, rk, "Acme", name="acme-org", out_dir=d)
    tls = {}
    for n in ("alpha", "beta", "gamma"):
        c, k, t = certutil.make_node(oc, ok, n, sans=("127.0.0.1",), out_dir=d)
        tls[n] = {"cert_pem": c, "key_pem": k, "trust_pem": t}
    crl = certutil.make_crl(oc, ok, (tls["beta"]["cert_pem"],), out_dir=d)
    tls["alpha"]["crls_pem"] = crl

    from hypha_b200.transport.utils import peer_id_from_key_pem as pid

    async def main():
        ids = {n: pid(tls[n]["key_pem"]) for n in tls}
        a = Network(ids["alpha"], tls=tls["alpha"])
        b = Network(ids["beta"], tls=tls["beta"])
        g = Network(ids["gamma"], tls=tls["gamma"])
        addr = await a.listen("tcp://127.0.0.1:0")
        assert await g.dial(addr) == ids["alpha"]                 # same CA, not revoked
        with pytest.raises(DialError):
            await b.dial(addr, timeout=2.0)
        assert ids["beta"] not in a.links and ids["gamma"] in a.links
        for n in (a, b, g):
            await n.close()
    asyncio.run(main())


def test_unconsumed_stream_does_not_stall_the_link(tmp_path):
    """Flow control is per stream (credits), not per link: a push stream that nobody reads stops after its
    window, while requests, pings and a second stream on the SAME socket link keep flowing.  (Round-1 defect:
    the link's read loop awaited the stream's bounded queue, so one ignored stream wedged the peer's link.)"""
    import hypha_b200.transport.network as N

    async def main():
        a, b = N.Network("a"), N.Network("b")
        addr = await a.listen(f"unix://{tmp_path}/a.sock")
        await b.dial(addr)

        async def echo(peer, msg):
            return {"echo": msg}

        a.on("/test/echo", None, echo)
        ignored = a.streams_push("job-ignored")        # registered, never consumed
        wanted = a.streams_push("job-wanted")
        big = bytes(N.STREAM_CHUNK) * (N.STREAM_WINDOW + 40)         # more than one window
        stuck = asyncio.ensure_future(b.stream_push(a.peer_id, big, job_id="job-ignored"))
        await asyncio.sleep(0.5)
        assert not stuck.done()                          # out of credit after STREAM_WINDOW chunks
        for i in range(5):                               # the link itself is alive
            assert await asyncio.wait_for(b.request(a.peer_id, "/test/echo", {"i": i}), 2.0) == {"echo": {"i": i}}
        assert await asyncio.wait_for(b.ping(a.peer_id), 2.0) >= 0
        payload = os.urandom(3 * N.STREAM_CHUNK + 17)
        push = asyncio.ensure_future(b.stream_push(a.peer_id, payload, job_id="job-wanted"))
        st = await asyncio.wait_for(wanted.get(), 2.0)
        assert await asyncio.wait_for(st.read_all(), 5.0) == payload and await push == len(payload)
        # the consumer finally looks at the ignored stream and refuses it: the sender gives up
        st2 = await asyncio.wait_for(ignored.get(), 2.0)
        st2.discard()
        stuck.cancel()
        for n in (a, b):
            await n.close()

    asyncio.run(main())


def test_gossip_is_signed_and_forgeries_are_dropped(tmp_path):
    """Gossip carries an Ed25519 signature of (topic, id, origin, payload) by the origin's key (the reference runs
    gossipsub with MessageAuthenticity::Signed).  A relaying peer that alters the payload, or publishes in
    somebody else's name, is ignored by nodes with a PKI; honest messages still arrive through the relay."""
    from hypha_b200 import certutil
    from hypha_b200.transport.network import Network
    from hypha_b200.transport.utils import peer_id_from_key_pem as pid

    d = str(tmp_path)
    rc, rk = certutil.make_root("Acme", out_dir=d)
    oc, ok = certutil.make_org(rc, rk, "Acme", name="acme-org", out_dir=d)
    tls = {}
    for n in ("gw", "sched", "worker", "mallory"):
        c, k, t = certutil.make_node(oc, ok, n, sans=("127.0.0.1",), out_dir=d)
        tls[n] = {"cert_pem": c, "key_pem": k, "trust_pem": t}
    ids = {n: pid(v["key_pem"]) for n, v in tls.items()}

    async def main():
        gw = Network(ids["gw"], tls=tls["gw"], relay_server=True, dht_server=True)
        addr = await gw.listen("tcp://127.0.0.1:0")
        sched, worker, mal = (Network(ids[n], tls=tls[n]) for n in ("sched", "worker", "mallory"))
        for n in (sched, worker, mal):
            await n.dial(addr)
        sub = await worker.subscribe("hypha/worker")
        await asyncio.sleep(0.2)
        assert await sched.publish("hypha/worker", b"ad-1") >= 1
        assert await sub.get(timeout=2.0) == (ids["sched"], b"ad-1")
        # mallory (a legitimate member of the PKI) forges an ad in the scheduler's name ...
        link = mal.links[ids["gw"]]
        forged = {"t": "pub", "topic": "hypha/worker", "mid": "f" * 32, "origin": ids["sched"], "data": b"ad-forged",
                  "proto": "/meshsub/1.1.0", "key": mal._pub_raw}
        forged["sig"] = mal._sign_key.sign(Network._gossip_digest(forged))
        await link.send(forged)
        # ... replays a signed ad with another payload, and sends an unsigned one
        await link.send({"t": "pub", "topic": "hypha/worker", "mid": "e" * 32, "origin": ids["mallory"], "data": b"unsigned",
                         "proto": "/meshsub/1.1.0"})
        await asyncio.sleep(0.4)
        assert gw.gossip_rejections == 2 and sub.q.empty()
        assert await mal.publish("hypha/worker", b"honest") >= 1            # her own signed messages are fine
        assert await sub.get(timeout=2.0) == (ids["mallory"], b"honest")
        for n in (gw, sched, worker, mal):
            await n.close()

    asyncio.run(main())
