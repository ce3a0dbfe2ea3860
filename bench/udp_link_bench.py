"""Bulk rate of one `udp://` link through one event loop (loopback), at the internet segment size and after path-MTU
discovery; tcp:// next to it for scale.  `python bench/udp_link_bench.py [--mb 64]` prints one JSON line.

The reference quotes 50-60 MB/s for its first libp2p stack and ~1 GB/s after tuning
(rfc/2025-03-25-libp2p_network_stack.md:9,17,29); the fabric's bulk path on a B200 box is peer memory, this link
carries the control plane and, across nodes, parameter-server files."""
import argparse
import asyncio
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from hypha_b200.transport.network import Network      # noqa: E402


async def one(scheme: str, nbytes: int, small_mss: bool) -> dict:
    a, b = Network("a"), Network("b")
    addr = await b.listen(f"{scheme}://127.0.0.1:0")
    await a.dial(addr)
    link = a.links["b"]
    if scheme == "udp":
        if small_mss:
            if link._probe_task is not None:
                link._probe_task.cancel()
            link._shrink_mss()
        else:
            await asyncio.sleep(1.5)                   # let path-MTU discovery finish
    q = b.streams_push()
    blob = os.urandom(nbytes)
    t0 = time.perf_counter()
    push = asyncio.ensure_future(a.stream_push("b", blob))
    st = await q.get()
    got = await st.read_all()
    await push
    dt = time.perf_counter() - t0
    assert len(got) == nbytes
    out = {"MBps": round(nbytes / dt / 1e6, 1)}
    if scheme == "udp":
        out.update(mss=link.mss, retransmits=link.retransmits)
    await a.close()
    await b.close()
    return out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=64)
    args = ap.parse_args()
    n = args.mb << 20
    res = {"bytes": n,
           "udp_mss1150": asyncio.run(one("udp", n, True)),
           "udp_pmtu": asyncio.run(one("udp", n, False)),
           "tcp": asyncio.run(one("tcp", n, False))}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
