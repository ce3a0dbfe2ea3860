"""`udp://` links (the reference's QUIC listener), direct-connection upgrade through the relay (dcutr) and the
Kademlia routing table / record replication (crates/network/src/kad.rs; Cargo.toml:48,54,106)."""
import asyncio
import os
import tempfile

import pytest

from hypha_b200.transport.kad import RESERVED_CIDRS, CidrFilter, RoutingTable, key_id
from hypha_b200.transport.network import DialError, Network, RecordNotFound


def run(coro):
    return asyncio.run(coro)


async def _echo_node(name, **kw):
    n = Network(name, **kw)

    async def echo(peer, m):
        return {"echo": m, "from": peer}

    n.on("/echo", None, echo)
    return n


def test_udp_link_request_and_lossy_stream():
    async def main():
        a, b = await _echo_node("a"), await _echo_node("b")
        addr = await b.listen("udp://127.0.0.1:0")
        assert addr.startswith("udp://127.0.0.1:") and not addr.endswith(":0")
        assert await a.dial(addr) == "b"
        assert (await a.request("b", "/echo", {"x": [1, 2, 3]})) == {"echo": {"x": [1, 2, 3]}, "from": "a"}
        assert (await b.request("a", "/echo", "back")) == {"echo": "back", "from": "b"}     # same connection, other way
        # a 3 MiB tensor stream with one datagram in seven dropped 40 times: every byte arrives, in order
        q = b.streams_push()
        blob = os.urandom(3 << 20)
        ep = a._udp[0]
        a.links["b"]._shrink_mss()                       # internet-sized segments (loopback would carry 60 KB ones)
        if a.links["b"]._probe_task is not None:
            a.links["b"]._probe_task.cancel()
        ep.drop_next, ep.drop_every = 40, 7
        push = asyncio.ensure_future(a.stream_push("b", blob))
        st = await asyncio.wait_for(q.get(), 20)
        got = await asyncio.wait_for(st.read_all(), 30)
        await push
        assert got == blob
        link = a.links["b"]
        assert 40 <= link.retransmits < 400 and ep.drop_next == 0
        assert sum(a.counters.outbound.values()) > len(blob)            # outbound bytes are accounted like on tcp links
        for n in (a, b):
            await n.close()

    run(main())


async def _udp_pair():
    from hypha_b200.transport.udplink import UdpEndpoint

    accepted = asyncio.get_event_loop().create_future()
    sep = UdpEndpoint(None, lambda link: accepted.set_result(link))
    sep.open("127.0.0.1", 0)
    cep = UdpEndpoint(None, lambda link: None)
    cep.open("127.0.0.1", 0)
    c = await cep.connect(*sep.local_addr)
    return sep, cep, await asyncio.wait_for(accepted, 5), c


def test_udp_receive_window_is_bounded_and_never_wedges():
    """Flow control of a udp link: (a) a frame larger than the receive window still arrives (the frame being assembled
    does not count against the window), (b) frames nobody reads stop the SENDER at about one window instead of piling
    up in the receiver, (c) when the reader returns the window update restarts the flow -- even when that update is
    lost (zero-window probe)."""
    from hypha_b200.transport import udplink

    async def main():
        sep, cep, s, c = await _udp_pair()
        big = os.urandom(udplink.RWND + (2 << 20))
        await asyncio.wait_for(c.send({"t": "stream_data", "sid": "x", "data": big}), 30)
        frame, _ = await asyncio.wait_for(s.recv(), 30)
        assert bytes(frame["data"]) == big
        chunk, n = os.urandom(256 << 10), 48                               # 12 MiB, three windows

        async def pour():
            for i in range(n):
                await c.send({"t": "stream_data", "sid": "y", "i": i, "data": chunk})

        task = asyncio.ensure_future(pour())
        await asyncio.sleep(1.5)
        assert not task.done()                                              # the sender is waiting for window
        assert s._queued <= udplink.RWND + 2 * len(chunk) and s._queued >= udplink.RWND // 2
        held = s._queued
        await asyncio.sleep(0.5)
        assert s._queued <= held + 64 * 1024                                # and it stays put (probes trickle at most)
        sep.drop_next, sep.drop_every = 4, 1                                # the window updates get lost
        for i in range(n):
            frame, _ = await asyncio.wait_for(s.recv(), 20)
            assert frame["i"] == i and bytes(frame["data"]) == chunk
        await asyncio.wait_for(task, 20)
        assert s._queued == 0 and sep.drop_next == 0
        cep.close()
        sep.close()

    run(main())


def test_udp_busy_server_asks_for_a_retry_cookie():
    """With many half-open connections an INIT is answered by a 33-byte RETRY carrying a cookie bound to the source
    address; the key exchange and the signature are spent only on an INIT that echoes it.  A client goes through this
    transparently; a flood from spoofed addresses costs the server one HMAC per datagram."""
    import socket
    import struct

    from hypha_b200.transport import udplink

    async def main():
        accepted = []
        sep = udplink.UdpEndpoint(None, accepted.append)
        sep.open("127.0.0.1", 0)
        # a flood: distinct (spoofable) client ids, no cookie
        raw = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
        raw.bind(("127.0.0.1", 0))
        raw.setblocking(False)
        for i in range(udplink.COOKIE_AFTER + 8):
            init = struct.pack("<B8s8s", udplink.INIT, b"\0" * 8, struct.pack("<Q", i + 1)) + os.urandom(32)
            raw.sendto(init + b"\0" * (udplink.INIT_MIN - len(init)), sep.local_addr)
            await asyncio.sleep(0.002)
        await asyncio.sleep(0.2)
        assert len(sep._pending_init) == udplink.COOKIE_AFTER and sep.retries_sent == 8
        kinds = []
        while True:
            try:
                kinds.append(raw.recv(65535))
            except BlockingIOError:
                break
        assert sum(1 for d in kinds if d[0] == udplink.RETRY and len(d) == 9 + udplink.COOKIE_LEN) == 8
        # a cookie issued to another address (or made up) is not accepted
        init = struct.pack("<B8s8s", udplink.INIT, b"\0" * 8, b"spoofer!") + os.urandom(32) + os.urandom(udplink.COOKIE_LEN)
        raw.sendto(init + b"\0" * (udplink.INIT_MIN - len(init)), sep.local_addr)
        await asyncio.sleep(0.1)
        assert len(sep._pending_init) == udplink.COOKIE_AFTER and sep.retries_sent == 9
        # a real client still gets in, one round trip later
        cep = udplink.UdpEndpoint(None, lambda link: None)
        cep.open("127.0.0.1", 0)
        c = await asyncio.wait_for(cep.connect(*sep.local_addr), 5)
        await asyncio.sleep(0.1)
        assert sep.retries_sent == 10 and len(accepted) == 1 and "cookie" in c.hs
        await c.send({"t": "hello"})
        assert (await asyncio.wait_for(accepted[0].recv(), 5))[0] == {"t": "hello"}
        raw.close()
        cep.close()
        sep.close()

    run(main())


def _pki(names):
    from hypha_b200 import certutil

    d = tempfile.mkdtemp(prefix="hbu", dir="/tmp")
    rc, rk = certutil.make_root("acme", out_dir=d)
    oc, ok = certutil.make_org(rc, rk, "acme", out_dir=d)
    out = {}
    for n in names:
        c, k, t = certutil.make_node(oc, ok, n, out_dir=d)
        out[n] = {"cert_pem": c, "key_pem": k, "trust_pem": t}
    return d, oc, ok, out


def test_udp_link_mutual_authentication_and_identity_binding():
    from hypha_b200 import certutil
    from hypha_b200.transport.utils import peer_id_from_key_pem

    async def main():
        d, oc, ok, tls = _pki(["n1", "n2", "n3"])
        ids = {n: peer_id_from_key_pem(t["key_pem"]) for n, t in tls.items()}
        n2 = await _echo_node(ids["n2"], tls=tls["n2"])
        addr = await n2.listen("udp://127.0.0.1:0")
        n1 = await _echo_node(ids["n1"], tls=tls["n1"])
        assert await n1.dial(addr) == ids["n2"]
        assert n1.links[ids["n2"]].cert_peer_id() == ids["n2"] and n2.links[ids["n1"]].cert_peer_id() == ids["n1"]
        assert (await n1.request(ids["n2"], "/echo", 7))["from"] == ids["n1"]
        # a CA-issued certificate entitles its holder to ITS key's peer id only
        imp = Network(ids["n2"], tls=tls["n1"])
        with pytest.raises(DialError):
            await imp.dial(addr, timeout=2.0)
        # no certificate: refused during the handshake, quickly
        anon = Network("anon")
        t0 = asyncio.get_event_loop().time()
        with pytest.raises(DialError):
            await anon.dial(addr, timeout=5.0)
        assert asyncio.get_event_loop().time() - t0 < 2.0
        # a certificate from another PKI
        _, _, _, other = _pki(["x"])
        stranger = Network(peer_id_from_key_pem(other["x"]["key_pem"]), tls=other["x"])
        with pytest.raises(DialError):
            await stranger.dial(addr, timeout=2.0)
        # a revoked node
        crl = certutil.make_crl(oc, ok, revoked=(tls["n3"]["cert_pem"],), out_dir=d)
        strict = await _echo_node(ids["n2"], tls=dict(tls["n2"], crls_pem=crl))
        saddr = await strict.listen("udp://127.0.0.1:0")
        n3 = Network(ids["n3"], tls=tls["n3"])
        with pytest.raises(DialError):
            await n3.dial(saddr, timeout=2.0)
        assert await n1.dial(saddr) == ids["n2"]             # the others are still welcome
        for n in (n1, n2, imp, anon, stranger, strict, n3):
            await n.close()

    run(main())


def test_dcutr_replaces_the_relayed_path_by_a_direct_udp_link():
    async def main():
        gw = Network("gw", relay_server=True, dht_server=True)
        gaddr = await gw.listen("tcp://127.0.0.1:0")
        a, b = await _echo_node("a"), await _echo_node("b")
        for n in (a, b):
            await n.listen("udp://127.0.0.1:0")
            n._udp[0].nat_filter = True       # unsolicited handshakes are dropped, like behind an address-restricted NAT
            await n.dial(gaddr)
        # behind the NAT a plain dial does not get through ...
        with pytest.raises(DialError):
            await a.dial(b.listen_addrs[0], timeout=0.5)
        a.peer_addrs.pop("b", None)
        b.peer_addrs.pop("a", None)
        a._udp[0]._contacted.clear()
        # ... so traffic goes through the gateway
        a._dcutr_last["b"] = 1e18             # keep the automatic upgrade out of this part of the test
        assert (await a.request("b", "/echo", 1))["from"] == "a"
        assert gw.relayed_frames >= 2 and "b" not in a.links
        # the upgrade: addresses swapped through the relay, both sides punch and dial
        assert await a.upgrade_direct("b")
        assert type(a.links["b"]).__name__ == "UdpLink" and b._udp[0].punches + a._udp[0].punches >= 1
        before = gw.relayed_frames
        assert (await a.request("b", "/echo", 2))["echo"] == 2
        assert (await b.request("a", "/echo", 3))["echo"] == 3
        assert gw.relayed_frames == before                   # the gateway is out of the data path
        # a peer that cannot be reached at all: the upgrade reports failure and the relay keeps working
        c = await _echo_node("c")
        await c.dial(gaddr)                                  # no listener at all
        a._dcutr_last["c"] = 1e18
        assert not await a.upgrade_direct("c", timeout=1.0)
        assert (await a.request("c", "/echo", 4))["echo"] == 4
        for n in (a, b, c, gw):
            await n.close()

    run(main())


def test_routing_table_buckets_cidr_filter_and_closest():
    t = RoutingTable("me", k=2)
    peers = [f"p{i}" for i in range(200)]
    for p in peers:
        t.add(p, [f"tcp://8.8.{len(p)}.{int(p[1:]) % 250}:1"])
    assert 2 <= len(t) <= 2 * 256 and all(len(b) <= 2 for b in t.buckets)
    # full buckets prefer their old, live entries; a stale entry makes room
    bi = next(i for i, b in enumerate(t.buckets) if len(b) == 2)
    old = t.buckets[bi]
    newcomer = next(p for p in (f"q{i}" for i in range(100000)) if t._t.bucket_index(p) == bi and p not in old)
    assert not t.add(newcomer, ["tcp://8.8.8.8:1"]) and t.buckets[bi] == old
    t.mark_disconnected(old[0])
    assert t.add(newcomer, ["tcp://8.8.8.8:1"]) and old[0] not in t.buckets[bi] and newcomer in t.buckets[bi]
    # refreshing an entry moves it to the most-recently-seen end and keeps its addresses when none are given
    first = t.buckets[bi][0]
    assert t.add(first, []) and t.buckets[bi][-1] == first and t.addrs(first)
    # the bucket index is the length of the id prefix shared with our own id
    import hashlib

    def lz(p):
        d = int.from_bytes(hashlib.sha256(p.encode()).digest(), "big") ^ int.from_bytes(hashlib.sha256(b"me").digest(), "big")
        return 256 - d.bit_length()

    assert all(t._t.bucket_index(p) == lz(p) for p in peers[:50])
    # closest() is the XOR order
    want = sorted(t.peers(), key=lambda p: key_id(p) ^ key_id("some-key"))[:5]
    assert t.closest("some-key", 5) == want
    # strict mode: reserved ranges never enter the table, an allow-list re-admits a site's own range
    strict = RoutingTable("me", cidr=CidrFilter(RESERVED_CIDRS, allow=("10.1.0.0/16",)))
    assert not strict.add("lo", ["tcp://127.0.0.1:1", "udp://192.168.1.5:2"]) and strict.rejected_by_cidr == 1
    assert strict.add("pub", ["tcp://127.0.0.1:1", "tcp://93.184.216.34:1"]) and strict.addrs("pub") == ["tcp://93.184.216.34:1"]
    assert strict.add("site", ["udp://10.1.2.3:9"]) and not strict.add("site2", ["udp://10.2.2.3:9"])
    assert strict.add("v6", ["tcp://[2a00:1450::1]:1"]) and not strict.add("v6lo", ["tcp://[::1]:1"])
    assert strict.add("dns", ["tcp://gateway.example.org:1"])


def test_two_gateways_replicate_records_and_lookups_walk_the_table():
    async def main():
        g1 = Network("g1", relay_server=True, dht_server=True)
        g2 = Network("g2", relay_server=True, dht_server=True)
        for g in (g1, g2):
            g.enable_dht_handlers()
        a1 = await g1.listen("tcp://127.0.0.1:0")
        a2 = await g2.listen("tcp://127.0.0.1:0")
        await g1.dial(a2)                                    # the gateways know each other
        assert "g2" in g1.kad and "g1" in g2.kad
        data, sched = Network("data"), Network("sched")
        await data.dial(a1)                                  # the data node only knows gateway 1 ...
        await sched.dial(a2)                                 # ... the scheduler only gateway 2
        # the scheduler learnt about gateway 1 from gateway 2's identify (other DHT servers it routes to); a record
        # only gateway 1 holds is found by walking there: no link to g1 existed before the lookup
        assert "g1" not in sched.links and "g1" in sched.kad and sched.kad.addrs("g1") == [a1]
        g1._records["only-g1"] = (b"v", "data")
        assert (await sched.get_record("only-g1")) == (b"v", "data")
        assert "g1" in sched.links
        await data.put_record("dataset", b'{"num_slices": 4}')
        await data.provide("dataset")
        await asyncio.sleep(0.2)
        assert g1._records["dataset"][1] == "data" and g2._records["dataset"] == g1._records["dataset"]
        assert (await sched.get_record("dataset")) == (b'{"num_slices": 4}', "data")
        assert await sched.find_providers("dataset") == ["data"]
        with pytest.raises(RecordNotFound):
            await sched.get_record("nobody-has-this", timeout=1.0)
        assert set(await sched.find_closest_peers("whatever")) >= {"g1", "g2"}
        # replication frames are taken from DHT servers only
        before = g2.identity_rejections
        await sched._send(sched.links["g2"], {"t": "dht_replicate", "key": "forged", "value": b"x", "publisher": "data"}, "/ipfs/kad/1.0.0")
        await asyncio.sleep(0.1)
        assert "forged" not in g2._records and g2.identity_rejections == before + 1
        for n in (data, sched, g1, g2):
            await n.close()

    run(main())


def test_dht_maintenance_republishes_to_a_gateway_that_joined_later():
    async def main():
        g1 = Network("g1", relay_server=True, dht_server=True)
        g1.enable_dht_handlers()
        a1 = await g1.listen("tcp://127.0.0.1:0")
        data = Network("data")
        await data.dial(a1)
        await data.put_record("dataset", b"v1")
        await data.provide("dataset")
        # a second gateway appears and meets the first one; the data node has never heard of it
        g2 = Network("g2", relay_server=True, dht_server=True)
        g2.enable_dht_handlers()
        a2 = await g2.listen("tcp://127.0.0.1:0")
        await g2.dial(a1)
        assert "dataset" not in g2._records and "g2" not in data.kad
        out = await data.dht_maintain_once()
        await asyncio.sleep(0.2)
        assert out["republished"] == 1 and out["servers_learnt"] >= 1 and "g2" in data.kad
        # own id + one id inside each of the far buckets (bucket refresh)
        keys = data.kad.refresh_keys()
        assert out["lookups"] == 1 + len(keys) and [data.kad._t.bucket_index(k) for k in keys] == list(range(len(keys)))
        assert len(keys) == len(set(keys)) >= 8 and data.kad.refresh_keys() == keys
        assert g2._records["dataset"] == (b"v1", "data") and "data" in g2._providers["dataset"]
        # the timer variant
        await data.put_record("dataset", b"v2")
        g2._records.clear()
        t = data.start_dht_maintenance(0.1)
        await asyncio.sleep(0.5)
        assert g2._records["dataset"] == (b"v2", "data")
        t.cancel()
        for n in (data, g1, g2):
            await n.close()

    run(main())


def test_replicated_records_are_verifiable_under_a_pki():
    """With certificates every record carries its publisher's Ed25519 signature: a DHT server cannot plant a record
    under somebody else's name through the server-to-server replication path."""
    from hypha_b200.transport.utils import peer_id_from_key_pem

    async def main():
        _, _, _, tls = _pki(["g1", "g2", "data", "rogue"])
        ids = {n: peer_id_from_key_pem(t["key_pem"]) for n, t in tls.items()}
        g1 = Network(ids["g1"], relay_server=True, dht_server=True, tls=tls["g1"])
        g2 = Network(ids["g2"], relay_server=True, dht_server=True, tls=tls["g2"])
        for g in (g1, g2):
            g.enable_dht_handlers()
        a1 = await g1.listen("udp://127.0.0.1:0")
        a2 = await g2.listen("udp://127.0.0.1:0")
        await g1.dial(a2)
        data = Network(ids["data"], tls=tls["data"])
        await data.dial(a1)
        await data.put_record("dataset", b"real")
        await asyncio.sleep(0.3)
        assert g2._records["dataset"] == (b"real", ids["data"])          # replicated, signature verified
        # a rogue DHT server (valid certificate of its own) forges a record in the data node's name
        rogue = Network(ids["rogue"], dht_server=True, tls=tls["rogue"])
        await rogue.dial(a2)
        before = g2.identity_rejections
        forged = {"t": "dht_replicate", "key": "dataset", "value": b"poison", "publisher": ids["data"], "kind": "record"}
        await rogue._send(rogue.links[ids["g2"]], dict(forged), "/ipfs/kad/1.0.0")                       # unsigned
        await rogue._send(rogue.links[ids["g2"]], rogue._sign_record(dict(forged)) | {"publisher": ids["data"]}, "/ipfs/kad/1.0.0")
        await asyncio.sleep(0.2)
        assert g2._records["dataset"] == (b"real", ids["data"]) and g2.identity_rejections == before + 2
        for n in (data, rogue, g1, g2):
            await n.close()

    run(main())


def test_record_store_limits_lifetime_ownership_and_order():
    """`kad.RecordStore`: bounded (value / key size, per-publisher and per-node counts, providers per key), expiring,
    a live key belongs to its publisher when ownership is enforced, and an older sequence number does not replace a
    newer one."""
    from hypha_b200.transport import kad

    now = [0.0]
    st = kad.RecordStore(ttl_s=100.0, enforce_owner=True, clock=lambda: now[0])
    assert st.put("k", b"v1", "alice", seq=5) == "stored" and st["k"] == (b"v1", "alice") and st["k"].seq == 5
    assert st.put("k", b"old", "alice", seq=4) == "stale" and st["k"][0] == b"v1"          # replayed older copy
    assert st.put("k", b"v2", "alice", seq=6) == "stored" and st["k"][0] == b"v2"
    assert st.put("k", b"mine now", "mallory", seq=99) == "owned" and st["k"] == (b"v2", "alice")
    now[0] = 101.0                                                                           # nobody republished it
    assert "k" not in st and st.get("k") is None
    assert st.put("k", b"m", "mallory", seq=1) == "stored"                                   # the key is free again
    assert st.put("big", b"x" * (kad.MAX_RECORD_VALUE + 1), "alice") == "too-large"
    assert st.put("k" * 300, b"", "alice") == "too-large"
    for i in range(kad.MAX_RECORDS_PER_PUBLISHER):
        assert st.put(f"bulk-{i}", b"", "spammer") == "stored"
    assert st.put("one-more", b"", "spammer") == "full" and st.put("other", b"", "alice") == "stored"
    assert st.put("bulk-0", b"again", "spammer", seq=1) == "stored"                          # refreshing is not growth
    now[0] = 250.0
    assert st.put("one-more", b"", "spammer") == "stored" and len(st) == 1                   # expired entries made room
    assert st.put("mine", b"", "me", local=True) == "stored"
    now[0] = 1e9
    assert "mine" in st and "one-more" not in st                                             # own records stay
    # providers
    for i in range(kad.MAX_PROVIDERS_PER_KEY):
        assert st.add_provider("file", f"p{i}") == "stored"
    assert st.add_provider("file", "p-extra") == "full" and st.add_provider("file", "p0") == "stored"
    assert len(st.providers("file")) == kad.MAX_PROVIDERS_PER_KEY
    now[0] += 101.0
    assert st.providers("file") == set() and st.add_provider("file", "p-extra") == "stored"
    assert st.refused["owned"] == 1 and st.refused["stale"] == 1 and st.refused["full"] >= 2
    # without ownership (no PKI: peer ids are per process, a restarted data node must be able to publish again)
    loose = kad.RecordStore(clock=lambda: 0.0)
    assert loose.put("k", b"a", "x") == "stored" and loose.put("k", b"b", "y") == "stored" and loose["k"] == (b"b", "y")


def test_pki_dht_keys_are_owned_and_reads_are_verified():
    """Under a PKI (a) another certificate holder cannot take over a live key -- the dataset record keeps pointing at
    the data node that published it -- and (b) a reader checks the publisher's signature on the copy a server hands
    out, so a rogue DHT server cannot answer a lookup with a record of its own making."""
    from hypha_b200.transport.utils import peer_id_from_key_pem

    async def main():
        _, _, _, tls = _pki(["g1", "data", "mallory", "rogue", "sched"])
        ids = {n: peer_id_from_key_pem(t["key_pem"]) for n, t in tls.items()}
        g1 = Network(ids["g1"], relay_server=True, dht_server=True, tls=tls["g1"])
        g1.enable_dht_handlers()
        a1 = await g1.listen("tcp://127.0.0.1:0")
        data = Network(ids["data"], tls=tls["data"])
        await data.dial(a1)
        await data.put_record("dataset", b"real")
        mallory = Network(ids["mallory"], tls=tls["mallory"])
        await mallory.dial(a1)
        await mallory.put_record("dataset", b"mine")              # accepted by the link, refused by the store
        await asyncio.sleep(0.2)
        assert g1._records["dataset"] == (b"real", ids["data"]) and g1.record_refusals == 1
        await data.put_record("dataset", b"real-v2")              # the owner itself can update it
        await asyncio.sleep(0.2)
        assert g1._records["dataset"] == (b"real-v2", ids["data"])
        # an old signed copy replayed by its publisher's link (or a server that kept it) does not roll v2 back
        old = data._sign_record({"t": "dht_put", "key": "dataset", "value": b"real", "publisher": ids["data"]})
        old["seq"] = 1
        old["sig"] = data._sign_key.sign(data._record_digest("record", "dataset", b"real", ids["data"], 1))
        await data._send(data.links[ids["g1"]], old, "/ipfs/kad/1.0.0")
        await asyncio.sleep(0.2)
        assert g1._records["dataset"][0] == b"real-v2"
        # a scheduler reads through an honest server: verified
        sched = Network(ids["sched"], tls=tls["sched"])
        await sched.dial(a1)
        assert await sched.get_record("dataset") == (b"real-v2", ids["data"])
        # ... and through a rogue server that answers with a forged record: refused, the lookup goes on and fails
        rogue = Network(ids["rogue"], relay_server=True, dht_server=True, tls=tls["rogue"])
        rogue.enable_dht_handlers()
        ar = await rogue.listen("tcp://127.0.0.1:0")
        rogue._records["dataset"] = (b"poison", ids["data"])      # no signature of the data node exists for this
        s2 = Network(ids["sched"], tls=tls["sched"])
        await s2.dial(ar)
        before = s2.identity_rejections
        with pytest.raises(RecordNotFound):
            await s2.get_record("dataset", timeout=1.0)
        assert s2.identity_rejections == before + 1
        for n in (data, mallory, sched, s2, rogue, g1):
            await n.close()

    run(main())


def test_observed_address_makes_a_wildcard_listener_reachable_for_hole_punching():
    """A node that binds udp://0.0.0.0 does not know an address others can use; the gateway tells it how its
    datagrams arrive (identify's observed address), and the direct-connection upgrade works with that."""
    async def main():
        gw = Network("gw", relay_server=True, dht_server=True)
        gaddr = await gw.listen("udp://127.0.0.1:0")
        a, b = await _echo_node("a"), await _echo_node("b")
        for n in (a, b):
            await n.listen("udp://0.0.0.0:0")
            n._udp[0].nat_filter = True
            await n.dial(gaddr)                       # over the node's one UDP socket
        port_b = b._udp[0].local_addr[1]
        assert b.observed_addrs == [f"udp://127.0.0.1:{port_b}"] and a.observed_addrs
        assert b._dcutr_addrs() == [f"udp://127.0.0.1:{port_b}"]          # the wildcard bind is not advertised
        a.peer_addrs.pop("b", None)
        b.peer_addrs.pop("a", None)
        a._dcutr_last["b"] = 1e18
        assert (await a.request("b", "/echo", 1))["from"] == "a" and "b" not in a.links      # relayed
        assert await a.upgrade_direct("b")
        before = gw.relayed_frames
        assert (await a.request("b", "/echo", 2))["echo"] == 2 and gw.relayed_frames == before
        for n in (a, b, gw):
            await n.close()

    run(main())


def test_udp_handshake_gives_no_amplification_and_fits_small_mtus():
    """An INIT shorter than 1200 bytes is not answered (a spoofed source cannot turn 49 bytes into a kilobyte of
    certificates aimed at a victim), and with a three-level PKI the handshake datagrams stay under 1280 bytes."""
    import socket
    import struct

    from hypha_b200.transport import udplink
    from hypha_b200.transport.utils import peer_id_from_key_pem

    async def main():
        _, _, _, tls = _pki(["s", "c"])
        srv = Network(peer_id_from_key_pem(tls["s"]["key_pem"]), tls=tls["s"])
        addr = await srv.listen("udp://127.0.0.1:0")
        port = int(addr.rsplit(":", 1)[1])
        probe = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
        probe.setblocking(False)
        probe.bind(("127.0.0.1", 0))
        short = struct.pack("<B8s8s", udplink.INIT, b"\0" * 8, b"12345678") + b"\1" * 32
        probe.sendto(short, ("127.0.0.1", port))
        await asyncio.sleep(0.2)
        with pytest.raises(BlockingIOError):
            probe.recvfrom(4096)                                  # silence
        probe.sendto(short + b"\0" * (udplink.INIT_MIN - len(short)), ("127.0.0.1", port))
        await asyncio.sleep(0.2)
        resp, _ = probe.recvfrom(4096)
        assert resp[0] == udplink.RESP and len(resp) <= 1280 and len(resp) <= udplink.INIT_MIN + 80
        probe.close()
        assert len(srv._udp[0]._chain) == 2                       # node + organisation CA; the root is not sent
        cli = Network(peer_id_from_key_pem(tls["c"]["key_pem"]), tls=tls["c"])
        assert await cli.dial(addr) == srv.peer_id                # and the real handshake still verifies the chain
        for n in (cli, srv):
            await n.close()

    run(main())


def test_addresses_learnt_from_peers_pass_the_cidr_filter():
    """What a peer says about itself (or, as a relay, about others) is dialled later: with `exclude_cidr` set such
    addresses are dropped, so a peer cannot point this node at a range the operator ruled out."""
    async def main():
        gw = Network("gw", relay_server=True, dht_server=True)
        gaddr = await gw.listen("tcp://127.0.0.1:0")
        liar = Network("liar")
        liar.add_external_address("tcp://10.0.0.7:22")
        liar.add_external_address("tcp://93.184.216.34:4001")
        await liar.dial(gaddr)
        strict = Network("strict", exclude_cidrs=("10.0.0.0/8",))
        await strict.dial(gaddr)
        assert strict.peer_addrs["liar"] == ["tcp://93.184.216.34:4001"]          # via the gateway's peer table
        assert gw.peer_addrs["liar"] == ["tcp://10.0.0.7:22", "tcp://93.184.216.34:4001"]     # no filter configured
        for n in (liar, strict, gw):
            await n.close()

    run(main())


def test_udp_path_mtu_discovery_and_fallback():
    """Segments grow to what the path carries (probes with DF set, acknowledged sizes only) and fall back to the
    conservative size when large datagrams stop getting through."""
    async def main():
        a, b = Network("a"), Network("b")
        addr = await b.listen("udp://127.0.0.1:0")
        await a.listen("udp://127.0.0.1:0")
        a._udp[0].mtu_limit = b._udp[0].mtu_limit = 9000          # a jumbo-frame fabric
        await a.dial(addr)
        la, lb = a.links["b"], b.links["a"]
        for _ in range(100):                                        # discovery runs in the background
            if la.mss == lb.mss == 8900 - 44 and la._probe_task.done() and lb._probe_task.done():
                break
            await asyncio.sleep(0.05)
        assert la.mss == 8900 - 44 and lb.mss == 8900 - 44          # 8.9 KB probe acknowledged, 16 KB refused
        q = b.streams_push()
        blob = os.urandom(4 << 20)
        push = asyncio.ensure_future(a.stream_push("b", blob))
        st = await asyncio.wait_for(q.get(), 10)
        assert await asyncio.wait_for(st.read_all(), 20) == blob
        await push
        sent_before = la._pn
        assert sent_before < 1200                                   # ~500 data packets instead of ~3700 at 1150 bytes
        # the route changes to a 1500-byte path while a transfer is on its way
        a._udp[0].mtu_limit = b._udp[0].mtu_limit = 1500
        blob2 = os.urandom(1 << 20)
        push = asyncio.ensure_future(a.stream_push("b", blob2))
        st = await asyncio.wait_for(q.get(), 20)
        assert await asyncio.wait_for(st.read_all(), 40) == blob2
        await push
        assert la.mss == 1150 and la.retransmits > 0
        for n in (a, b):
            await n.close()

    run(main())


def test_udp_reassembly_survives_any_arrival_order():
    """Property: whatever order, duplication and re-cutting the segments of a stream arrive in, the receiver hands out
    exactly the frames that were sent, and its out-of-order store ends empty."""
    import struct

    from hypothesis import given, settings, strategies as st

    from hypha_b200.transport import udplink

    class FakeEndpoint:
        def __init__(self):
            self._dirty, self.conns, self.sent = set(), {}, 0

        def sendto(self, data, addr):
            self.sent += 1
            return True

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.binary(min_size=0, max_size=700), min_size=1, max_size=6), st.data())
    def prop(payloads, data):
        async def main():
            import cbor2

            ep = FakeEndpoint()
            link = udplink.UdpLink(ep, ("127.0.0.1", 1), b"12345678", False)
            link._tx_key = type("K", (), {"encrypt": staticmethod(lambda n, p, a: p + b"\0" * 16)})()
            frames = [{"t": "x", "i": i, "data": p} for i, p in enumerate(payloads)]
            stream = b"".join(struct.pack("<I", len(b)) + b for b in (cbor2.dumps(f) for f in frames))
            # cut the stream twice with different segment sizes (the sender re-cuts what is in flight when its
            # segment size drops), shuffle everything, repeat some
            segs = []
            for size in (data.draw(st.integers(1, 400)), data.draw(st.integers(1, 400))):
                segs += [(o, stream[o:o + size]) for o in range(0, len(stream), size)]
            order = data.draw(st.permutations(segs))
            order = list(order) + data.draw(st.lists(st.sampled_from(segs), max_size=8))
            first = [s for s in order]
            # make sure every byte arrives at least once in the first cut's terms: the full list contains both cuts
            for off, seg in first:
                link._on_segment(off, seg)
            got = []
            while link._frames:
                got.append((await link.recv())[0])
            assert [g["i"] for g in got] == list(range(len(frames)))
            assert [bytes(g["data"]) for g in got] == payloads
            assert link._rcv_nxt == len(stream) and len(link._stream) == 0 and link._queued == 0
            link._in_order = 255                                   # the periodic sweep of stale parked segments
            link._ooo[0] = b"stale"
            link._ooo_bytes += 5
            link._on_segment(link._rcv_nxt, b"")
            assert all(o >= link._rcv_nxt for o in link._ooo) or not link._ooo
            if link._timer is not None:
                link._timer.cancel()

        asyncio.run(main())

    prop()


def test_record_store_invariants_under_random_operations():
    """Property: whatever sequence of puts, provider announcements and clock jumps a store sees, it stays within its
    bounds, its per-publisher accounting matches its contents, an owner's record is never replaced by somebody else
    while it is alive, and a publisher's sequence numbers never go backwards."""
    from hypothesis import given, settings, strategies as st

    from hypha_b200.transport import kad

    ops = st.lists(st.one_of(
        st.tuples(st.just("put"), st.integers(0, 5), st.integers(0, 3), st.integers(0, 50), st.booleans()),
        st.tuples(st.just("prov"), st.integers(0, 5), st.integers(0, 3)),
        st.tuples(st.just("tick"), st.floats(0.0, 80.0))), max_size=60)

    @settings(max_examples=80, deadline=None)
    @given(ops)
    def prop(seq):
        now = [0.0]
        s = kad.RecordStore(ttl_s=100.0, enforce_owner=True, clock=lambda: now[0])
        for op in seq:
            if op[0] == "put":
                _, k, p, sq, local = op
                key, pub = f"k{k}", f"p{p}"
                before = s.get(key)
                res = s.put(key, b"v%d" % sq, pub, seq=sq, local=local)
                after = s.get(key)
                if before is not None and before.publisher != pub and not local:
                    assert res == "owned" and after == before
                if before is not None and before.publisher == pub and sq < before.seq:
                    assert res == "stale" and after.seq == before.seq
                if res == "stored":
                    assert after == (b"v%d" % sq, pub) and after.seq == sq
            elif op[0] == "prov":
                s.add_provider(f"k{op[1]}", f"p{op[2]}")
            else:
                now[0] += op[1]
            live = [s.get(k) for k in s.keys()]
            live = [r for r in live if r is not None]
            assert len(s) <= kad.MAX_RECORDS
            counts = {}
            for r in live:
                counts[r.publisher] = counts.get(r.publisher, 0) + 1
            for p, n in counts.items():
                assert s._per_pub.get(p, 0) >= n          # expired-but-unswept entries may still be counted
            s.expire()
            assert {p: n for p, n in s._per_pub.items()} == counts
            assert all(len(s.providers(f"k{k}")) <= kad.MAX_PROVIDERS_PER_KEY for k in range(6))

    prop()
