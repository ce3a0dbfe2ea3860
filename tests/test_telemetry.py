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


def test_gpu_monitor_samples_exports_and_raises_the_alarm():
    """Worker resource monitor against a fake NVML: samples become hypha.gpu.* gauges with a gpu attribute, power-cap
    throttling is normal, a hardware / thermal slowdown is reported once per change."""
    import types

    from opentelemetry.sdk.metrics.export import InMemoryMetricReader

    from hypha_b200.telemetry.gpu_monitor import GpuMonitor, decode_reasons

    state = {"mask": 0x4, "temp": 61}

    class FakeNvml:
        NVML_TEMPERATURE_GPU, NVML_CLOCK_SM = 0, 1
        calls = []

        def nvmlInit(self): self.calls.append("init")
        def nvmlShutdown(self): self.calls.append("shutdown")
        def nvmlDeviceGetCount(self): return 2
        def nvmlDeviceGetHandleByIndex(self, i): return i
        def nvmlDeviceGetUtilizationRates(self, h): return types.SimpleNamespace(gpu=97 - h, memory=40)
        def nvmlDeviceGetMemoryInfo(self, h): return types.SimpleNamespace(used=(100 + h) << 30, total=180 << 30)
        def nvmlDeviceGetTemperature(self, h, kind): return state["temp"] + h
        def nvmlDeviceGetPowerUsage(self, h): return 995_000
        def nvmlDeviceGetClockInfo(self, h, kind): return 1455
        def nvmlDeviceGetCurrentClocksEventReasons(self, h): return state["mask"] if h == 1 else 0x4

    assert decode_reasons(0x4 | 0x40) == ["sw_power_cap", "hw_thermal_slowdown"]
    logs = []
    reader = InMemoryMetricReader()
    tel = T.Telemetry("gpu", metric_reader=reader)
    mon = GpuMonitor(nvml=FakeNvml(), log=logs.append)
    assert mon.devices == [0, 1]
    mon.register(tel)
    mon._tick()
    s = mon.last()
    assert s[0]["utilization_pct"] == 97 and s[1]["memory_used"] == 101 << 30 and s[0]["power_w"] == 995.0
    assert s[0]["throttle_reasons"] == ["sw_power_cap"] and mon.health() == {"healthy": True, "throttled": {}} and not logs
    data = reader.get_metrics_data()
    got = {}
    for rm in data.resource_metrics:
        for sm in rm.scope_metrics:
            for m in sm.metrics:
                if m.name.startswith("hypha.gpu."):
                    got[m.name] = {dp.attributes["gpu"]: dp.value for dp in m.data.data_points}
    assert got["hypha.gpu.utilization"] == {"0": 97.0, "1": 96.0} and got["hypha.gpu.clock.sm"]["1"] == 1455.0
    assert got["hypha.gpu.temperature"] == {"0": 61.0, "1": 62.0} and got["hypha.gpu.throttled"] == {"0": 0, "1": 0}
    # GPU 1 overheats: one log line, the gauge flips, health says which device and why; then it recovers
    state["mask"], state["temp"] = 0x4 | 0x40 | 0x8, 91
    mon._tick()
    mon._tick()
    assert len(logs) == 1 and "gpu 1" in logs[0] and "hw_thermal_slowdown" in logs[0] and "hw_slowdown" in logs[0]
    assert mon.health() == {"healthy": False, "throttled": {1: ["hw_slowdown", "hw_thermal_slowdown"]}}
    state["mask"] = 0x4
    mon._tick()
    assert len(logs) == 2 and "normal clocks" in logs[1] and mon.health()["healthy"]
    mon.stop()
    assert "shutdown" in FakeNvml.calls
    # a machine without NVML: the worker simply has no monitor
    from hypha_b200.telemetry.gpu_monitor import start_for_worker

    cfg = types.SimpleNamespace(resources=types.SimpleNamespace(gpu=0.0), device=None)
    assert start_for_worker(cfg, tel, logs.append) is None


def test_training_metrics_become_otel_gauges():
    """With an OTLP endpoint the scheduler's metrics bridge also records `hypha.job.<metric>` gauges (worker and round
    as attributes); a sink that fails does not keep the others from getting the metric."""
    import asyncio

    from opentelemetry.sdk.metrics.export import InMemoryMetricReader

    from hypha_b200.scheduler import FanOutConnector, MemoryConnector, OtelConnector
    from hypha_b200.scheduler.scheduling import Connector

    class Broken(Connector):
        async def send(self, *a):
            raise RuntimeError("sink down")

    reader = InMemoryMetricReader()
    tel = T.Telemetry("sched", metric_reader=reader)
    mem = MemoryConnector()
    sink = FanOutConnector(Broken(), mem, OtelConnector(tel))

    async def main():
        await sink.send("worker-a", 1, "loss", 2.5)
        await sink.send("worker-b", 1, "loss", 2.75)
        await sink.send("worker-a", 2, "loss", 2.0)
        await sink.send("worker-a", 2, "samples_or_tokens_per_s", 21940.0)
        await sink.send("worker-a", 2, "odd name/with spaces", 1.0)

    asyncio.run(main())
    got = {}
    for rm in reader.get_metrics_data().resource_metrics:
        for sm in rm.scope_metrics:
            for m in sm.metrics:
                for p in m.data.data_points:
                    if m.name.startswith("hypha.job."):
                        got[(m.name, p.attributes["worker"], p.attributes["round"])] = p.value
    assert got[("hypha.job.loss", "worker-a", 1)] == 2.5 and got[("hypha.job.loss", "worker-b", 1)] == 2.75
    assert got[("hypha.job.loss", "worker-a", 2)] == 2.0
    assert got[("hypha.job.samples_or_tokens_per_s", "worker-a", 2)] == 21940.0
    assert ("hypha.job.odd_name_with_spaces", "worker-a", 2) in got
    assert len(mem.records) == 5
    tel.shutdown()
