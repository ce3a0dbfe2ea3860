"""``udp://`` links: an encrypted, mutually authenticated, reliable datagram transport -- the role QUIC plays in the
reference (libp2p ``quic`` feature, Cargo.toml:48,106; every role listens on ``/ip4/.../udp/.../quic-v1`` next to tcp,
crates/scheduler/src/network.rs:119,145) and the substrate of hole punching (`Network.upgrade_direct`, the
reference's dcutr): ONE socket per node carries every connection, so the NAT mapping a peer learns through the relay
is the one a direct dial lands on.

Not the QUIC wire format (no aioquic / TLS-over-UDP stack in this image) but the same construction:

* handshake, three datagrams (SIGMA / Noise-XX shape): ephemeral X25519 both ways; each side then proves the identity
  of its CA-issued certificate with an Ed25519 signature over the transcript, sent under the handshake key so
  certificates never travel in clear.  The peer id of the link is derived from the certificate key exactly as on the
  mTLS tcp links (`utils.peer_id_from_cert_der`), the chain is verified against ``trust_pem`` (+ CRLs);
* record protection: ChaCha20-Poly1305 per direction, keys from HKDF-SHA256 over the DH secret and the transcript
  hash, nonce = packet number, header authenticated;
* reliability: one ordered byte stream per connection (the fabric multiplexes its own streams on top, with its own
  credit-based flow control): segments <= 1150 bytes, cumulative + duplicate ACKs, RTT-based retransmission timer with
  back-off, fast retransmit on 3 duplicate ACKs, a fixed congestion window, receiver window advertised in every ACK;
* the byte stream carries the same length-prefixed CBOR frames as tcp / unix links (`network.SocketLink`).
"""
from __future__ import annotations

import asyncio
import collections
import datetime as dt
import os
import struct
import time
from typing import Dict, List, Optional, Tuple

import cbor2

INIT, RESP, FIN, DATA, CLOSE, PUNCH, RETRY = 1, 2, 3, 4, 5, 6, 7
COOKIE_LEN = 24                 # 8-byte timestamp + 16 bytes of HMAC-SHA256 over (source address, client cid, timestamp)
COOKIE_AFTER = 64               # half-open connections beyond which an INIT must echo a cookie (QUIC's Retry)
COOKIE_TTL_S = 10.0
MSS = 1150                      # stream bytes per datagram (fits a 1280-byte path MTU with headers + tag)
CWND = 1024 * 1024              # upper bound of the congestion window (bytes in flight)
INIT_CWND = 16 * 1150
INIT_MIN = 1200                 # smallest INIT datagram a server answers
# Path-MTU discovery (the socket sets DF): after the handshake each side probes larger datagrams and raises its
# segment size to the largest one the peer acknowledged -- 1150 bytes suit the internet, a jumbo-frame fabric carries
# 8.9 KB, loopback 60 KB, and the per-datagram cost (crypto, syscall, bookkeeping) is what bounds the stream rate.
PROBE_SIZES = (1400, 4000, 8900, 16000, 32000, 60000)
OVERHEAD = 17 + 16 + 11         # packet header + AEAD tag + segment record header
RWND = 4 * 1024 * 1024          # reassembly buffer the receiver advertises
MIN_RTO, MAX_RTO = 0.05, 2.0
MAX_FRAME = 16 << 20
SPLIT_BIT = 0x80000000


class UdpError(Exception):
    pass


def _hkdf(secret: bytes, salt: bytes, info: bytes, n: int) -> bytes:
    from cryptography.hazmat.primitives import hashes
    from cryptography.hazmat.primitives.kdf.hkdf import HKDF

    return HKDF(algorithm=hashes.SHA256(), length=n, salt=salt, info=info).derive(secret)


def verify_peer_chain(chain_der: List[bytes], trust_pem: str, crls_pem: Optional[str] = None) -> bytes:
    """Verify ``chain_der`` (leaf first) up to a certificate of ``trust_pem``; returns the leaf's raw Ed25519 public
    key.  Checks signatures, validity periods, the CA flag of every issuer and (optionally) revocation of the leaf."""
    from cryptography import x509
    from cryptography.hazmat.primitives import serialization

    if not chain_der:
        raise UdpError("peer sent no certificate")
    chain = [x509.load_der_x509_certificate(d) for d in chain_der]
    roots = x509.load_pem_x509_certificates(open(trust_pem, "rb").read())
    now = dt.datetime.now(dt.timezone.utc)

    def valid_now(c) -> bool:
        return c.not_valid_before_utc <= now <= c.not_valid_after_utc

    def is_ca(c) -> bool:
        try:
            return bool(c.extensions.get_extension_for_class(x509.BasicConstraints).value.ca)
        except x509.ExtensionNotFound:
            return False

    def fp(c) -> bytes:
        return c.fingerprint(_sha256())

    root_fps = {fp(r) for r in roots}
    cur, pool, hops = chain[0], chain[1:], 0
    if not valid_now(cur):
        raise UdpError("peer certificate is not valid at this time")
    while fp(cur) not in root_fps:
        hops += 1
        if hops > 8:
            raise UdpError("certificate chain too long")
        issuer = None
        for c in list(roots) + pool:
            try:
                cur.verify_directly_issued_by(c)
                issuer = c
                break
            except Exception:
                continue
        if issuer is None or not is_ca(issuer) or not valid_now(issuer):
            raise UdpError("peer certificate does not chain to a trusted root")
        pool = [c for c in pool if c is not issuer]
        cur = issuer
    if crls_pem:
        data = open(crls_pem, "rb").read()
        blocks = [b"-----BEGIN X509 CRL-----" + p for p in data.split(b"-----BEGIN X509 CRL-----")[1:]]
        for blk in blocks:
            crl = x509.load_pem_x509_crl(blk)
            if crl.issuer == chain[0].issuer and crl.get_revoked_certificate_by_serial_number(chain[0].serial_number):
                raise UdpError("peer certificate is revoked")
    pub = chain[0].public_key()
    try:
        return pub.public_bytes(serialization.Encoding.Raw, serialization.PublicFormat.Raw)
    except ValueError as exc:
        raise UdpError("peer certificate key is not Ed25519") from exc


def _sha256():
    from cryptography.hazmat.primitives import hashes

    return hashes.SHA256()


class UdpLink:
    """One connection of a `UdpEndpoint`: `send(frame)` / `recv()` with the interface of `network.SocketLink`."""

    def __init__(self, ep: "UdpEndpoint", addr, scid: bytes, initiator: bool):
        self.ep, self.addr, self.scid, self.initiator = ep, addr, scid, initiator
        self.dcid = b"\0" * 8
        self.peer_id: Optional[str] = None
        self.closed = False
        self.established = asyncio.get_event_loop().create_future()
        self._cert_peer: Optional[str] = None
        self._tx_key = self._rx_key = None
        self._pn = 0
        self._rx_seen: set = set()
        self._rx_floor = 0
        # sender
        self._snd_buf = bytearray()                 # not yet segmented
        self._snd_nxt = 0                           # next stream offset to assign
        self._snd_una = 0
        self._unacked: "collections.OrderedDict[int, list]" = collections.OrderedDict()   # off -> [data, t_sent, retx]
        self._inflight = 0
        self._peer_wnd = RWND
        self._dup_acks = 0
        self._recover = -1                          # NewReno: stream offset that ends the current loss recovery
        self._cwnd = INIT_CWND                      # slow start / congestion avoidance / halving on loss
        self._ssthresh = CWND
        self._srtt: Optional[float] = None
        self._rttvar = 0.0
        self._rto = 0.2
        self._space = asyncio.Event()
        self._space.set()
        self._wlock = asyncio.Lock()
        # receiver
        self._rcv_nxt = 0
        self._ooo: Dict[int, bytes] = {}
        self._ooo_bytes = 0
        self._stream = bytearray()
        self._frames: collections.deque = collections.deque()
        self._queued = 0                            # bytes of complete frames nobody has `recv()`-ed yet
        self._adv_wnd = RWND                        # the window last advertised
        self._in_order = 0
        self._persist, self._persist_at = 0.2, 0.0  # zero-window probing (sender side)
        self._frame_waiter: Optional[asyncio.Future] = None
        self._ack_due = False
        self._segs_since_ack = 0
        self._timer: Optional[asyncio.TimerHandle] = None
        self.retransmits = 0
        self.mss = MSS                                  # raised by `_probe_mtu`, dropped back by repeated time-outs
        self._probe_acked = 0
        self._probe_task: Optional[asyncio.Task] = None
        self.hs: dict = {}

    # -- identity ---------------------------------------------------------------------------------------------------
    def cert_peer_id(self) -> Optional[str]:
        return self._cert_peer

    # -- packets ------------------------------------------------------------------------------------------------------
    def _seal(self, payload: bytes) -> bytes:
        hdr = struct.pack("<B8sQ", DATA, self.dcid, self._pn)
        ct = self._tx_key.encrypt(b"\0\0\0\0" + struct.pack("<Q", self._pn), payload, hdr)
        self._pn += 1
        return hdr + ct

    def _open(self, data: bytes) -> Optional[bytes]:
        if self._rx_key is None or len(data) < 17 + 16:
            return None
        (pn,) = struct.unpack_from("<Q", data, 9)
        if pn < self._rx_floor or pn in self._rx_seen:
            return None                                           # replay
        try:
            pt = self._rx_key.decrypt(b"\0\0\0\0" + data[9:17], data[17:], data[:17])
        except Exception:
            return None
        self._rx_seen.add(pn)
        if len(self._rx_seen) > 4096:
            self._rx_floor = max(self._rx_seen) - 2048
            self._rx_seen = {p for p in self._rx_seen if p >= self._rx_floor}
        return pt

    def _send_records(self, recs: bytes) -> None:
        if not self.closed or recs[:1] == b"C":
            self.ep.sendto(self._seal(recs), self.addr)

    def flush_ack(self) -> None:
        if self._ack_due and not self.closed and self._tx_key is not None:
            self._ack_due = False
            self._send_records(self._ack_record())

    def _window(self) -> int:
        """What the receiver can still take: the reassembly budget minus what is parked out of order and minus the
        complete frames the application has not read yet.  The frame under assembly does not count -- it is bounded by
        `MAX_FRAME` on its own, and counting it would close the window on any frame larger than the window."""
        return max(0, RWND - self._ooo_bytes - self._queued)

    def _ack_record(self) -> bytes:
        self._segs_since_ack = 0
        self._adv_wnd = self._window()
        return b"A" + struct.pack("<QI", self._rcv_nxt, self._adv_wnd)

    # -- sending -------------------------------------------------------------------------------------------------------
    def _pump(self) -> None:
        now = time.monotonic()
        while self._snd_buf and self._inflight < min(self._cwnd, self._peer_wnd):
            seg = bytes(self._snd_buf[:self.mss])
            del self._snd_buf[:self.mss]
            off = self._snd_nxt
            self._snd_nxt += len(seg)
            self._unacked[off] = [seg, now, 0]
            self._inflight += len(seg)
            rec = b"S" + struct.pack("<QH", off, len(seg)) + seg
            if self._ack_due:
                rec += self._ack_record()
                self._ack_due = False
            self._send_records(rec)
        if len(self._snd_buf) < CWND:                 # `send` parks at most ~2 windows here, whatever the peer's window
            self._space.set()
        else:
            self._space.clear()
        self._arm()

    def _arm(self) -> None:
        if self._timer is None and (self._unacked or self._ack_due or self._snd_buf) and not self.closed:
            self._timer = asyncio.get_event_loop().call_later(0.01 if self._ack_due else min(self._rto, 0.05), self._tick)

    def _tick(self) -> None:
        self._timer = None
        if self.closed:
            return
        now = time.monotonic()
        if self._ack_due:
            self._ack_due = False
            self._send_records(self._ack_record())
        if self._snd_buf and not self._unacked and self._peer_wnd < self.mss and now >= self._persist_at:
            # zero-window probe: the receiver's window update may have been lost, and nothing else is in flight whose
            # ACK would bring a new one.  One segment, at a doubling interval; its ACK carries the current window.
            seg = bytes(self._snd_buf[:self.mss])
            del self._snd_buf[:self.mss]
            off = self._snd_nxt
            self._snd_nxt += len(seg)
            self._unacked[off] = [seg, now, 0]
            self._inflight += len(seg)
            self._send_records(b"S" + struct.pack("<QH", off, len(seg)) + seg)
            self._persist = min(2.0, self._persist * 2)
            self._persist_at = now + self._persist
        elif self._unacked:
            off, ent = next(iter(self._unacked.items()))
            if now - ent[1] > self._rto:
                if ent[2] >= 12:
                    self._fail("peer unreachable")
                    return
                if ent[2] >= 2 and self.mss > MSS:
                    self._shrink_mss()                       # the path may have stopped carrying large datagrams
                rto = self._rto
                self._rto = min(MAX_RTO, self._rto * 2)
                self._recover = self._snd_nxt                # what follows the hole may be lost as well
                self._ssthresh = max(4 * MSS, self._cwnd // 2)
                self._cwnd = 4 * MSS
                budget = self._cwnd                          # go-back-N, one (collapsed) window per timeout
                for o, e in self._unacked.items():
                    if budget <= 0 or now - e[1] <= rto:
                        break
                    e[1], e[2] = now, e[2] + 1
                    self.retransmits += 1
                    budget -= len(e[0])
                    self._send_records(b"S" + struct.pack("<QH", o, len(e[0])) + e[0])
        self._arm()

    def _on_ack(self, ack: int, wnd: int) -> None:
        self._peer_wnd = wnd
        now = time.monotonic()
        if wnd >= self.mss:
            self._persist, self._persist_at = 0.2, 0.0
        elif self._persist_at == 0.0:
            self._persist_at = now + self._persist
        if ack > self._snd_una:
            self._snd_una = ack
            self._dup_acks = 0
            if self._srtt is not None:                       # progress: undo the timer back-off
                self._rto = min(MAX_RTO, max(MIN_RTO, self._srtt + 4 * self._rttvar))
            while self._unacked:
                off, ent = next(iter(self._unacked.items()))
                if off + len(ent[0]) > ack:
                    break
                self._unacked.popitem(last=False)
                self._inflight -= len(ent[0])
                if ack >= self._recover:                     # not while repairing a loss
                    grow = len(ent[0]) if self._cwnd < self._ssthresh else max(1, MSS * len(ent[0]) // self._cwnd)
                    self._cwnd = min(CWND, self._cwnd + grow)
                if ent[2] == 0:                              # Karn: only samples of segments sent once
                    r = now - ent[1]
                    if self._srtt is None:
                        self._srtt, self._rttvar = r, r / 2
                    else:
                        self._rttvar = 0.75 * self._rttvar + 0.25 * abs(self._srtt - r)
                        self._srtt = 0.875 * self._srtt + 0.125 * r
                    self._rto = min(MAX_RTO, max(MIN_RTO, self._srtt + 4 * self._rttvar))
            if ack < self._recover and self._unacked:        # partial ACK inside a recovery: the next hole, at once
                self._retransmit_first(now)
        elif ack == self._snd_una and self._unacked:
            self._dup_acks += 1
            if self._dup_acks == 3 and ack >= self._recover:
                self._recover = self._snd_nxt
                self._ssthresh = max(4 * MSS, self._cwnd // 2)
                self._cwnd = self._ssthresh
                self._retransmit_first(now)
        self._pump()

    def _shrink_mss(self) -> None:
        """Back to the conservative segment size; what is in flight is re-cut so that retransmissions fit."""
        self.mss = MSS
        cut: "collections.OrderedDict[int, list]" = collections.OrderedDict()
        for off, (data, t_sent, retx) in self._unacked.items():
            for i in range(0, len(data), MSS):
                cut[off + i] = [data[i:i + MSS], t_sent, retx]
        self._unacked = cut

    async def _probe_mtu(self) -> None:
        try:
            for size in PROBE_SIZES:
                ok = False
                for _ in range(2):
                    if self.closed:
                        return
                    rec = b"P" + struct.pack("<I", size)
                    rec += b"\0" * max(0, size - 17 - 16 - len(rec))
                    if not self.ep.sendto(self._seal(rec), self.addr):
                        break                                # EMSGSIZE: larger than the interface MTU
                    await asyncio.sleep(0.05 + 2 * (self._srtt or 0.05))
                    if self._probe_acked >= size:
                        ok = True
                        break
                if not ok:
                    return
                self.mss = max(self.mss, size - OVERHEAD)
        except Exception:               # noqa: BLE001 - discovery is an optimisation
            return

    def _retransmit_first(self, now: float) -> None:
        off, ent = next(iter(self._unacked.items()))
        ent[1], ent[2] = now, ent[2] + 1
        self.retransmits += 1
        self._send_records(b"S" + struct.pack("<QH", off, len(ent[0])) + ent[0])

    async def send(self, frame: dict) -> int:
        from .network import SPLIT_FRAME_MIN, NetworkError

        if self.closed:
            raise NetworkError("link closed")
        payload = frame.get("data") if frame.get("t") == "stream_data" else None
        if payload is not None and len(payload) >= SPLIT_FRAME_MIN:
            meta = cbor2.dumps({k: v for k, v in frame.items() if k != "data"})
            total = 4 + len(meta) + len(payload)
            parts = [struct.pack("<II", total | SPLIT_BIT, len(meta)), meta, bytes(payload)]
        else:
            data = cbor2.dumps(frame)
            total = len(data)
            parts = [struct.pack("<I", total), data]
        async with self._wlock:
            for p in parts:
                pos = 0
                while pos < len(p):
                    await self._space.wait()
                    if self.closed:
                        raise NetworkError("link closed")
                    n = min(len(p) - pos, CWND)
                    self._snd_buf += p[pos:pos + n]
                    pos += n
                    self._pump()
        return total

    # -- receiving -------------------------------------------------------------------------------------------------------
    def on_datagram(self, data: bytes) -> None:
        pt = self._open(data)
        if pt is None:
            return
        i, n = 0, len(pt)
        while i < n:
            k = pt[i:i + 1]
            if k == b"S" and i + 11 <= n:
                off, ln = struct.unpack_from("<QH", pt, i + 1)
                seg = pt[i + 11:i + 11 + ln]
                i += 11 + ln
                self._on_segment(off, seg)
            elif k == b"A" and i + 13 <= n:
                ack, wnd = struct.unpack_from("<QI", pt, i + 1)
                i += 13
                self._on_ack(ack, wnd)
            elif k == b"P" and i + 5 <= n:                  # MTU probe: the rest of the datagram is padding
                (size,) = struct.unpack_from("<I", pt, i + 1)
                self._send_records(b"Q" + struct.pack("<I", size))
                return
            elif k == b"Q" and i + 5 <= n:
                (size,) = struct.unpack_from("<I", pt, i + 1)
                self._probe_acked = max(self._probe_acked, int(size))
                i += 5
            elif k == b"C":
                self._fail("closed by peer", notify=False)
                return
            else:
                return

    def _on_segment(self, off: int, seg: bytes) -> None:
        self._ack_due = True
        if off == self._rcv_nxt:
            self._stream += seg
            self._rcv_nxt += len(seg)
            while self._rcv_nxt in self._ooo:
                s = self._ooo.pop(self._rcv_nxt)
                self._ooo_bytes -= len(s)
                self._stream += s
                self._rcv_nxt += len(s)
            self._in_order += 1
            if self._ooo and self._in_order % 256 == 0:
                # segments parked under offsets the stream has passed by other cuts (the sender re-cuts what is in
                # flight when it lowers its segment size) would hold window space for good
                for o in [o for o in self._ooo if o < self._rcv_nxt]:
                    self._ooo_bytes -= len(self._ooo.pop(o))
            self._parse()
            self._segs_since_ack += 1
            self.ep._dirty.add(self)                         # one ACK per batch of datagrams read (`_on_readable`)
        elif off > self._rcv_nxt:
            if off not in self._ooo and self._ooo_bytes + len(seg) <= RWND:
                self._ooo[off] = seg
                self._ooo_bytes += len(seg)
            # out of order: acknowledge at once so that the sender sees duplicate ACKs
            self._ack_due = False
            self._send_records(self._ack_record())
            return
        self._arm()

    def _parse(self) -> None:
        buf = self._stream
        while len(buf) >= 4:
            (n,) = struct.unpack_from("<I", buf, 0)
            split = bool(n & SPLIT_BIT)
            n &= ~SPLIT_BIT
            if n > MAX_FRAME:
                self._fail("frame too large")
                return
            if len(buf) < 4 + n:
                break
            with memoryview(buf) as mv:                      # one copy out of the stream buffer, not two
                self._frames.append((bytearray(mv[4:4 + n]), split))
            self._queued += n
            del buf[:4 + n]
        w, self._frame_waiter = self._frame_waiter, None
        if w is not None and not w.done():
            w.set_result(None)

    async def recv(self) -> Optional[Tuple[dict, int]]:
        from .network import NetworkError

        while not self._frames:
            if self.closed:
                return None
            self._frame_waiter = asyncio.get_event_loop().create_future()
            await self._frame_waiter
        body, split = self._frames.popleft()
        n, m = len(body), 0
        self._queued -= n
        if self._adv_wnd < RWND // 4 and self._window() >= RWND // 2 and not self.closed and self._tx_key is not None:
            self._send_records(self._ack_record())           # window update: the sender was (nearly) stopped
        if split:
            if n < 4:
                raise NetworkError("malformed bulk frame")
            (m,) = struct.unpack_from("<I", body, 0)
            if 4 + m > n:
                raise NetworkError("malformed bulk frame")
        try:
            frame = cbor2.loads(bytes(body[4:4 + m]) if split else bytes(body))
        except cbor2.CBORError as exc:
            raise NetworkError(f"undecodable frame: {exc}") from exc
        if not isinstance(frame, dict):
            raise NetworkError("frame is not a map")
        if split:
            frame["data"] = memoryview(body)[4 + m:]
        return frame, n

    # -- teardown ---------------------------------------------------------------------------------------------------------
    def _fail(self, why: str, notify: bool = True) -> None:
        if self.closed:
            return
        if notify and self._tx_key is not None:
            try:
                self.ep.sendto(self._seal(b"C"), self.addr)
            except Exception:
                pass
        self.closed = True
        self.ep.conns.pop(self.scid, None)
        if self._timer is not None:
            self._timer.cancel()
            self._timer = None
        if self._probe_task is not None:
            self._probe_task.cancel()
        self._space.set()
        if not self.established.done():
            self.established.set_exception(UdpError(why))
            if not self.initiator:
                self.established.exception()                 # nobody awaits a half-open server connection
        w, self._frame_waiter = self._frame_waiter, None
        if w is not None and not w.done():
            w.set_result(None)

    async def close(self) -> None:
        self._fail("closed")


class UdpEndpoint:
    """One UDP socket, many connections (demultiplexed by the 8-byte destination connection id)."""

    def __init__(self, tls: Optional[dict], on_accept):
        self.tls, self.on_accept = tls, on_accept
        self.sock = None
        self._dirty: set = set()                  # links that received in-order data in the current read batch
        self.conns: Dict[bytes, UdpLink] = {}
        self._pending_init: Dict[Tuple, UdpLink] = {}      # (addr, client cid) -> half-open server connection
        self.local_addr: Optional[Tuple[str, int]] = None
        self.punches = 0
        self.drop_next = 0            # test hook: drop this many outgoing DATA datagrams ...
        self.drop_every = 1           # ... one out of every `drop_every`
        self._drop_ctr = 0
        self.nat_filter = False       # test hook: behave like an address-restricted NAT (INITs only from punched peers)
        self.mtu_limit = 0            # test hook: refuse datagrams above this size, like an interface MTU
        self._contacted: set = set()
        self._cookie_secret = os.urandom(32)
        self._swept_at = 0.0
        self.retries_sent = 0
        self._key = self._chain = None
        if tls:
            from cryptography import x509
            from cryptography.hazmat.primitives import serialization

            self._key = serialization.load_pem_private_key(open(tls["key_pem"], "rb").read(), password=None)
            # leaf first; the self-signed root stays home (the peer holds it as its trust anchor), which keeps the
            # handshake datagrams of a three-level PKI under a 1280-byte MTU
            certs = x509.load_pem_x509_certificates(open(tls["cert_pem"], "rb").read())
            self._chain = [c.public_bytes(serialization.Encoding.DER) for i, c in enumerate(certs)
                           if i == 0 or c.issuer != c.subject]

    def open(self, host: str, port: int) -> None:
        """Bind the node's one UDP socket.  Reads are drained in batches (up to 64 datagrams per wake-up of the event
        loop instead of one callback per datagram -- the loop iteration was the largest per-packet cost), and every
        connection acknowledges once per batch."""
        import socket

        fam = socket.AF_INET6 if ":" in host else socket.AF_INET
        sock = socket.socket(fam, socket.SOCK_DGRAM)
        sock.setblocking(False)
        for opt in (socket.SO_RCVBUF, socket.SO_SNDBUF):         # as much as the kernel grants (rmem_max / wmem_max)
            try:
                sock.setsockopt(socket.SOL_SOCKET, opt, 4 << 20)
            except OSError:
                pass
        try:                                                      # DF: an oversized probe fails instead of fragmenting
            if fam == socket.AF_INET:
                sock.setsockopt(socket.IPPROTO_IP, getattr(socket, "IP_MTU_DISCOVER", 10), getattr(socket, "IP_PMTUDISC_DO", 2))
            else:
                sock.setsockopt(socket.IPPROTO_IPV6, getattr(socket, "IPV6_MTU_DISCOVER", 23), getattr(socket, "IPV6_PMTUDISC_DO", 2))
        except OSError:
            pass
        sock.bind((host, port))
        self.sock = sock
        self.local_addr = sock.getsockname()[:2]
        asyncio.get_running_loop().add_reader(sock.fileno(), self._on_readable)

    def _on_readable(self) -> None:
        sock = self.sock
        if sock is None:
            return
        for _ in range(64):
            try:
                data, addr = sock.recvfrom(65535)
            except (BlockingIOError, InterruptedError):
                break
            except OSError:
                break
            self.datagram_received(data, addr)
        dirty, self._dirty = self._dirty, set()
        for link in dirty:
            link.flush_ack()

    def sendto(self, data: bytes, addr) -> bool:
        """False when the datagram could not even be handed to the kernel (EMSGSIZE for an MTU probe, no socket)."""
        if self.sock is None:
            return False
        if self.drop_next > 0 and data[:1] == bytes([DATA]):
            self._drop_ctr += 1
            if self._drop_ctr % self.drop_every == 0:
                self.drop_next -= 1
                return True
        if self.mtu_limit and len(data) > self.mtu_limit:
            return False                           # test hook: an interface with a smaller MTU
        self._contacted.add(tuple(addr[:2]))
        try:
            self.sock.sendto(data, addr)
        except (BlockingIOError, InterruptedError):
            pass                                   # a full send buffer is a lost datagram: the sender retransmits
        except OSError:
            return False
        return True

    def close(self) -> None:
        for c in list(self.conns.values()):
            c._fail("endpoint closed")
        if self.sock is not None:
            try:
                asyncio.get_event_loop().remove_reader(self.sock.fileno())
            except Exception:
                pass
            self.sock.close()
            self.sock = None

    # -- handshake -------------------------------------------------------------------------------------------------------
    @staticmethod
    def _aead(key: bytes):
        from cryptography.hazmat.primitives.ciphers.aead import ChaCha20Poly1305

        return ChaCha20Poly1305(key)

    def _auth_blob(self, transcript: bytes, role: bytes) -> bytes:
        if self._key is None:
            return cbor2.dumps({"chain": [], "sig": b""})
        return cbor2.dumps({"chain": self._chain, "sig": self._key.sign(b"hypha-udp-" + role + transcript)})

    def _check_auth(self, blob: bytes, transcript: bytes, role: bytes) -> Optional[str]:
        """Peer id the other side proved, or None on an unauthenticated link; raises when authentication is
        required (this node runs with certificates) and missing or wrong."""
        from cryptography.hazmat.primitives.asymmetric.ed25519 import Ed25519PublicKey

        from .utils import peer_id_from_public_bytes

        msg = cbor2.loads(blob)
        chain, sig = msg.get("chain") or [], msg.get("sig") or b""
        if self.tls is None:
            return None                       # this node does not authenticate peers (like tcp without TLS)
        raw = verify_peer_chain(list(chain), self.tls["trust_pem"], self.tls.get("crls_pem"))
        try:
            Ed25519PublicKey.from_public_bytes(raw).verify(sig, b"hypha-udp-" + role + transcript)
        except Exception as exc:
            raise UdpError("peer does not hold the key of its certificate") from exc
        return peer_id_from_public_bytes(raw)

    @staticmethod
    def _keys(shared: bytes, transcript: bytes):
        from cryptography.hazmat.primitives import hashes

        h = hashes.Hash(hashes.SHA256())
        h.update(transcript)
        th = h.finalize()
        okm = _hkdf(shared, th, b"hypha-udp keys", 96)
        return okm[:32], okm[32:64], okm[64:]            # handshake key, client->server, server->client

    async def connect(self, host: str, port: int, timeout: float = 5.0) -> UdpLink:
        from cryptography.hazmat.primitives import serialization
        from cryptography.hazmat.primitives.asymmetric.x25519 import X25519PrivateKey

        scid = os.urandom(8)
        link = UdpLink(self, (host, port), scid, True)
        e = X25519PrivateKey.generate()
        e_pub = e.public_key().public_bytes(serialization.Encoding.Raw, serialization.PublicFormat.Raw)
        link.hs = {"e": e, "e_pub": e_pub}
        self.conns[scid] = link
        # padded to INIT_MIN bytes: the server answers an INIT with ~1 KB (certificates) before the client has shown
        # that it can receive at its source address, so the request must not be smaller than the answer (QUIC's rule
        # against reflection amplification)
        head = struct.pack("<B8s8s", INIT, b"\0" * 8, scid) + e_pub
        deadline = time.monotonic() + timeout
        delay = 0.1
        while not link.established.done():
            if "fin" in link.hs:
                self.sendto(link.hs["fin"], link.addr)            # our FIN may have been lost: the RESP came again
            else:
                # a busy server first answers RETRY with a cookie bound to our address; the INIT then carries it
                init = head + link.hs.get("cookie", b"\0" * COOKIE_LEN)
                self.sendto(init + b"\0" * max(0, INIT_MIN - len(init)), link.addr)
            left = deadline - time.monotonic()
            if left <= 0:
                link._fail("handshake timed out", notify=False)
                break
            try:
                await asyncio.wait_for(asyncio.shield(link.established), min(delay, left))
            except asyncio.TimeoutError:
                delay = min(delay * 2, 1.0)
            except Exception:
                break
        return await link.established

    def datagram_received(self, data: bytes, addr) -> None:
        if len(data) < 9:
            return
        typ, dcid = data[0], data[1:9]
        try:
            if typ == DATA:
                c = self.conns.get(dcid)
                if c is not None:
                    if c._rx_key is None:
                        if not c.initiator and "resp" in c.hs:
                            self.sendto(c.hs["resp"], addr)    # our RESP or the client's FIN was lost
                        return
                    before = len(c._rx_seen)
                    c.on_datagram(data)
                    if c.addr != addr and len(c._rx_seen) > before:
                        c.addr = addr               # NAT rebinding / migration: the packet authenticated itself
            elif typ == INIT:
                if self.nat_filter and tuple(addr[:2]) not in self._contacted:
                    return                          # a stateful NAT drops unsolicited inbound traffic
                self._on_init(data, addr)
            elif typ == RESP:
                self._on_resp(data, addr)
            elif typ == FIN:
                self._on_fin(data, addr)
            elif typ == CLOSE:
                c = self.conns.get(dcid)
                # only a handshake can be refused in clear: nothing authenticated has arrived on this connection yet
                if c is not None and (not c.established.done() or (c.initiator and not c._rx_seen)):
                    c._fail("refused by peer: " + data[9:].decode(errors="replace"), notify=False)
            elif typ == RETRY:
                c = self.conns.get(dcid)
                if c is not None and c.initiator and not c.established.done() and "fin" not in c.hs \
                        and len(data) >= 9 + COOKIE_LEN and "cookie" not in c.hs:
                    c.hs["cookie"] = data[9:9 + COOKIE_LEN]
                    init = struct.pack("<B8s8s", INIT, b"\0" * 8, c.scid) + c.hs["e_pub"] + c.hs["cookie"]
                    self.sendto(init + b"\0" * max(0, INIT_MIN - len(init)), c.addr)
            elif typ == PUNCH:
                self.punches += 1
        except UdpError as exc:
            c = self.conns.get(dcid)
            if c is not None:
                if c.dcid != b"\0" * 8:                      # tell the other side why instead of letting it time out
                    self.sendto(struct.pack("<B8s", CLOSE, c.dcid) + str(exc).encode()[:200], addr)
                c._fail(str(exc), notify=False)
        except Exception:
            return

    def _cookie(self, addr, ccid: bytes, ts: Optional[int] = None) -> bytes:
        import hashlib
        import hmac

        ts = int(time.time()) if ts is None else ts
        msg = repr(tuple(addr[:2])).encode() + ccid + struct.pack("<Q", ts)
        return struct.pack("<Q", ts) + hmac.new(self._cookie_secret, msg, hashlib.sha256).digest()[:16]

    def _cookie_ok(self, cookie: bytes, addr, ccid: bytes) -> bool:
        import hmac

        if len(cookie) != COOKIE_LEN:
            return False
        (ts,) = struct.unpack_from("<Q", cookie, 0)
        if abs(time.time() - ts) > COOKIE_TTL_S:
            return False
        return hmac.compare_digest(cookie, self._cookie(addr, ccid, ts))

    def _on_init(self, data: bytes, addr) -> None:
        from cryptography.hazmat.primitives import serialization
        from cryptography.hazmat.primitives.asymmetric.x25519 import X25519PrivateKey, X25519PublicKey

        if len(data) < INIT_MIN:
            return                                             # see `connect`: no amplification for spoofed sources
        ccid, e_c = data[9:17], data[17:49]
        half = self._pending_init.get((addr, ccid))
        if half is not None:                                   # retransmitted INIT: repeat our answer
            self.sendto(half.hs["resp"], addr)
            return
        # half-open connections older than 10 s are forgotten (an INIT flood must not pin memory)
        now = time.monotonic()
        if now - self._swept_at > 1.0:
            self._swept_at = now
            for k, v in list(self._pending_init.items()):
                if now - v.hs.get("t0", now) > 10.0:
                    self._pending_init.pop(k, None)
                    v._fail("handshake never completed", notify=False)
        if len(self._pending_init) >= COOKIE_AFTER and not self._cookie_ok(data[49:49 + COOKIE_LEN], addr, ccid):
            # many half-open connections: before spending a key exchange and a signature on this one, make the
            # client show that it receives at the address it claims (QUIC's Retry; 33 bytes back for 1200 in)
            self.retries_sent += 1
            self.sendto(struct.pack("<B8s", RETRY, ccid) + self._cookie(addr, ccid), addr)
            return
        now = time.monotonic()
        scid = os.urandom(8)
        link = UdpLink(self, addr, scid, False)
        link.dcid = ccid
        e = X25519PrivateKey.generate()
        e_pub = e.public_key().public_bytes(serialization.Encoding.Raw, serialization.PublicFormat.Raw)
        shared = e.exchange(X25519PublicKey.from_public_bytes(e_c))
        transcript = ccid + scid + e_c + e_pub
        hk, c2s, s2c = self._keys(shared, transcript)
        blob = self._aead(hk).encrypt(b"\0" * 11 + b"\1", self._auth_blob(transcript, b"server"), transcript)
        resp = struct.pack("<B8s8s", RESP, ccid, scid) + e_pub + blob
        link.hs = {"hk": hk, "transcript": transcript, "c2s": c2s, "s2c": s2c, "resp": resp, "t0": now}
        self.conns[scid] = link
        self._pending_init[(addr, ccid)] = link
        if len(self._pending_init) > 1024:                      # a flood: the oldest half-open connection goes
            k = next(iter(self._pending_init))
            self._pending_init.pop(k)._fail("too many half-open connections", notify=False)
        self.sendto(resp, addr)

    def _on_resp(self, data: bytes, addr) -> None:
        from cryptography.hazmat.primitives.asymmetric.x25519 import X25519PublicKey

        link = self.conns.get(data[1:9])
        if link is None or not link.initiator or len(data) < 17 + 32 + 16:
            return
        if "fin" in link.hs:
            self.sendto(link.hs["fin"], addr)
            return
        scid_s, e_s, blob = data[9:17], data[17:49], data[49:]
        shared = link.hs["e"].exchange(X25519PublicKey.from_public_bytes(e_s))
        transcript = link.scid + scid_s + link.hs["e_pub"] + e_s
        hk, c2s, s2c = self._keys(shared, transcript)
        try:
            auth = self._aead(hk).decrypt(b"\0" * 11 + b"\1", blob, transcript)
        except Exception as exc:
            raise UdpError("handshake authentication failed") from exc
        link._cert_peer = self._check_auth(auth, transcript, b"server")
        link.dcid = scid_s
        link.addr = addr
        fin_blob = self._aead(hk).encrypt(b"\0" * 11 + b"\2", self._auth_blob(transcript, b"client"), transcript)
        link.hs["fin"] = struct.pack("<B8s", FIN, scid_s) + fin_blob
        link._tx_key, link._rx_key = self._aead(c2s), self._aead(s2c)
        self.sendto(link.hs["fin"], addr)
        if not link.established.done():
            link.established.set_result(link)
            link._probe_task = asyncio.ensure_future(link._probe_mtu())

    def _on_fin(self, data: bytes, addr) -> None:
        link = self.conns.get(data[1:9])
        if link is None or link.initiator or link._tx_key is not None:
            return
        hs = link.hs
        try:
            auth = self._aead(hs["hk"]).decrypt(b"\0" * 11 + b"\2", data[9:], hs["transcript"])
        except Exception as exc:
            raise UdpError("handshake authentication failed") from exc
        link._cert_peer = self._check_auth(auth, hs["transcript"], b"client")
        link._tx_key, link._rx_key = self._aead(hs["s2c"]), self._aead(hs["c2s"])
        self._pending_init = {k: v for k, v in self._pending_init.items() if v is not link}
        if not link.established.done():
            link.established.set_result(link)
            link._probe_task = asyncio.ensure_future(link._probe_mtu())
        self.on_accept(link)

    def punch(self, host: str, port: int) -> None:
        """An unauthenticated datagram whose only purpose is to open this side's NAT mapping towards ``host:port``."""
        self.sendto(struct.pack("<B8s", PUNCH, b"\0" * 8), (host, port))
