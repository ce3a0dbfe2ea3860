"""The fabric in one file: three nodes in one process -- a gateway (relay + DHT server), a "server" and a "client"
-- exercising every capability a role is assembled from (the counterpart of the reference's
crates/network/examples/request_response):

* dial / listen on tcp:// (pass --scheme udp for the encrypted datagram links, memory for in-process links)
* typed request / response with pattern-matched handlers (the first matching registration wins)
* gossip: publish / subscribe
* DHT records: put on one node, get from another
* a push stream (bulk bytes) and a pull stream (header in, bytes out)

    python examples/fabric_request_response.py [--scheme tcp|udp|memory]
"""
import argparse
import asyncio
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from hypha_b200.transport.network import Network  # noqa: E402

ECHO = "/example-echo/0.0.1"


async def main(scheme: str) -> dict:
    gateway, server, client = Network("gateway", relay_server=True, dht_server=True), Network("server"), Network("client")
    for node in (gateway, server, client):
        node.enable_dht_handlers()                      # what `roles.start_node` does for every role
    where = "memory://gateway" if scheme == "memory" else f"{scheme}://127.0.0.1:0"
    gw_addr = await gateway.listen(where)
    for node in (server, client):                       # everybody bootstraps through the gateway
        await node.dial(gw_addr)
        await node.wait_for_bootstrap()
    srv_addr = await server.listen("memory://server" if scheme == "memory" else f"{scheme}://127.0.0.1:0")
    await client.dial(srv_addr)

    # -- request / response: two handlers on one protocol, chosen by a predicate on the message ----------------------
    async def shout(peer, msg):
        return {"echo": msg["text"].upper(), "by": "shout"}

    async def plain(peer, msg):
        return {"echo": msg["text"], "by": "plain"}

    server.on(ECHO, lambda m: m.get("loud"), shout)
    server.on(ECHO, None, plain)                        # matches everything the first one did not take
    loud = await client.request("server", ECHO, {"text": "hello", "loud": True})
    quiet = await client.request("server", ECHO, {"text": "hello"})

    # -- gossip ----------------------------------------------------------------------------------------------------
    sub = await server.subscribe("example/topic")
    await asyncio.sleep(0.1)                            # the subscription travels to the neighbours
    await client.publish("example/topic", b"an advertisement")
    origin, data = await sub.get(timeout=5)

    # -- DHT record: stored on the gateway (the DHT server), read by the other node -----------------------------------
    await server.put_record("datasets/example", b'{"num_slices": 3}')
    value, publisher = await client.get_record("datasets/example")

    # -- streams: push 1 MiB to the server; pull a "slice" from it ----------------------------------------------------
    incoming = server.streams_push()
    blob = os.urandom(1 << 20)
    push = asyncio.ensure_future(client.stream_push("server", blob))
    stream = await incoming.get()
    received = await stream.read_all()
    await push

    async def serve_slice(peer, header):
        return b"slice-%d-of-%s" % (header["index"], header["dataset"].encode())

    server.serve_pull(serve_slice)
    pulled = await (await client.stream_pull("server", {"dataset": "example", "index": 2})).read_all()

    out = {"loud": loud, "quiet": quiet, "gossip": [origin, data.decode()], "record": [value.decode(), publisher],
           "pushed_ok": received == blob, "pulled": pulled.decode(),
           "bytes_out_client": sum(client.counters.outbound.values())}
    for node in (client, server, gateway):
        await node.close()
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--scheme", choices=["tcp", "udp", "memory"], default="tcp")
    import json

    print(json.dumps(asyncio.run(main(ap.parse_args().scheme))))
