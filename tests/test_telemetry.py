"""Telemetry tests (model: crates/telemetry/src/bandwidth.rs:322-790 with the in-memory metric
exporter of testing.rs:8-45, and the proptest round-trips of otlp.rs / attributes.rs)."""
import json

import pytest
from hypothesis import given, strategies as st

from hypha_b200 import telemetry as T


def test_attribute_and_header_parsing():
    assert T.parse_kv_list("a=1, b=two%20words,c=") == {"a": "1", "b": "two words", "c": ""}
    assert T.parse_kv_list(None) == {} and T.parse_kv_list("") == {}
    with pytest.raises(T.OtlpConfigError):
        T.parse_kv_list("novalue")
    with pytest.raises(T.OtlpConfigError):
        T.parse_kv_list("=x")


@given(st.dictionaries(st.text(alphabet="abcdefghijklmnopqrstuvwxyz._-", min_size=1, max_size=8),
                       st.text(max_size=12), max_size=5))
def test_kv_roundtrip(d):
    assert T.parse_kv_list(T.format_kv_list(d)) == d


def test_protocol_and_endpoint():
    assert T.Protocol.parse(None).value == "grpc" and T.Protocol.parse("http/protobuf").value == "http/protobuf"
    with pytest.raises(T.OtlpConfigError):
        T.Protocol.parse("carrier-pigeon")
    assert T.Endpoint.parse(None) is None
    e = T.Endpoint.parse("http://collector:4318/")
    assert e.signal(T.Protocol.parse("http/protobuf"), "traces") == "http://collector:4318/v1/traces"
    assert e.signal(T.Protocol.parse("grpc"), "traces") == "http://collector:4318"
    with pytest.raises(T.OtlpConfigError):
        T.Endpoint.parse("collector:4317")


def test_sampler_selection():
    from opentelemetry.sdk.trace import sampling as s

    assert T.make_sampler(None, None) is s.ALWAYS_ON
    assert isinstance(T.make_sampler(None, 0.5), s.ParentBased)
    assert T.make_sampler("always_off", None) is s.ALWAYS_OFF
    assert isinstance(T.make_sampler("traceidratio", 0.1), s.TraceIdRatioBased)
    assert isinstance(T.make_sampler("parentbased_traceidratio", 0.1), s.ParentBased)
    with pytest.raises(T.OtlpConfigError):
        T.make_sampler("sometimes", None)


def test_noop_without_endpoint_and_jsonl_spans(tmp_path):
    path = tmp_path / "spans.jsonl"
    tel = T.Telemetry("test-role", jsonl_path=str(path))
    assert not tel.enabled
    with tel.span("outer-step", round=3):
        pass
    tel.shutdown()
    rec = json.loads(path.read_text().strip())
    assert rec["name"] == "outer-step" and rec["attributes"]["round"] == 3


def test_bandwidth_counters_emit_per_protocol_and_direction():
    """emits_inbound_bytes_on_read / does_not_pollute_outbound_on_read."""
    from opentelemetry.sdk.metrics.export import InMemoryMetricReader

    reader = InMemoryMetricReader()
    tel = T.Telemetry("bw", metric_reader=reader)
    c = T.BandwidthCounters()
    tel.register_bandwidth(c)
    c.add("/hypha-api/0.0.1", inbound=100)
    c.add("/hypha-api/0.0.1", inbound=23)
    c.add("/hypha-tensor-stream/push", outbound=4096)
    data = reader.get_metrics_data()
    got = {}
    for rm in data.resource_metrics:
        for sm in rm.scope_metrics:
            for m in sm.metrics:
                for p in m.data.data_points:
                    got[(m.name, p.attributes["protocols"])] = p.value
    assert got[(T.INBOUND, "/hypha-api/0.0.1")] == 123
    assert got[(T.OUTBOUND, "/hypha-tensor-stream/push")] == 4096
    assert (T.OUTBOUND, "/hypha-api/0.0.1") not in got
    assert c.total() == {T.INBOUND: 123, T.OUTBOUND: 4096}
    tel.shutdown()


def test_from_config_reads_telemetry_keys():
    from hypha_b200 import config as C

    cfg = C.builder(C.GatewayConfig).with_env({"OTEL_RESOURCE_ATTRIBUTES": "deployment=box7",
                                               "OTEL_TRACES_SAMPLER": "always_off"}).build()
    tel = T.Telemetry.from_config("gw", cfg)
    assert tel.resource.attributes["deployment"] == "box7" and tel.resource.attributes["service.name"] == "gw"
    tel.shutdown()


def test_otlp_http_export_reaches_a_collector():
    """With `telemetry_endpoint` set, traces, metrics (incl. the bandwidth counters) and logs are POSTed as
    OTLP/HTTP protobuf to `<endpoint>/v1/{traces,metrics,logs}` with the configured headers
    (telemetry/src/{tracing,metrics,logging}.rs).  The collector here is a local HTTP server."""
    import threading
    from http.server import BaseHTTPRequestHandler, HTTPServer

    from hypha_b200.telemetry import Telemetry
    from hypha_b200.telemetry.bandwidth import BandwidthCounters

    got = []

    class Collector(BaseHTTPRequestHandler):
        def do_POST(self):
            n = int(self.headers.get("Content-Length", 0))
            body = self.rfile.read(n)
            got.append((self.path, self.headers.get("Content-Type"), self.headers.get("x-api-key"), len(body)))
            self.send_response(200)
            self.send_header("Content-Type", "application/x-protobuf")
            self.end_headers()

        def log_message(self, *a):
            pass

    srv = HTTPServer(("127.0.0.1", 0), Collector)
    t = threading.Thread(target=srv.serve_forever, daemon=True)
    t.start()
    try:
        tel = Telemetry("hypha-test", endpoint=f"http://127.0.0.1:{srv.server_port}", protocol="http/protobuf",
                        headers="x-api-key=secret", attributes="deployment=ci", metric_interval_s=0.2)
        bw = BandwidthCounters()
        bw.add("/hypha-api/0.0.1", inbound=123, outbound=456)
        tel.register_bandwidth(bw)
        with tel.span("allocate.workers", num=2):
            tel.logger.warning("lease lost")
        import time

        time.sleep(0.5)                     # one periodic metric export
        tel.shutdown()                      # flushes the batch processors
    finally:
        srv.shutdown()
    paths = {p for p, *_ in got}
    assert {"/v1/traces", "/v1/metrics"} <= paths, got
    assert all(ct == "application/x-protobuf" and key == "secret" and size > 0 for p, ct, key, size in got
               if p in ("/v1/traces", "/v1/metrics"))
    assert "/v1/logs" in paths          # the log bridge exported the warning
