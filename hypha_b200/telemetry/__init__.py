"""Telemetry: OTLP traces / logs / metrics, JSONL spans, bandwidth counters, device timers.

Reference: crates/telemetry/src/lib.rs:15-50 (factories), tracing.rs:23-126 (samplers, exporters,
no-op layer), logging.rs:17-64, metrics.rs:13-75 (global meter, 1 s periodic exporter),
bandwidth.rs (byte counters).  GPU-side additions: NVTX ranges and CUDA-event timers,
because every reported multi-GPU number must be device-timed."""
from __future__ import annotations

import contextlib
import json
import logging
import os
import sys
import threading
import time
from typing import Any, Dict, Iterator, List, Optional

from .bandwidth import INBOUND, OUTBOUND, BandwidthCounters
from .otlp import Endpoint, OtlpConfigError, Protocol, format_kv_list, parse_kv_list

SAMPLERS = ("always_on", "always_off", "traceidratio", "parentbased_traceidratio")


def make_sampler(kind: Optional[str], ratio: Optional[float]):
    """Sampler selection (tracing.rs:23-32,60-80): explicit kind wins; a bare ratio means
    ParentBased(TraceIdRatio); nothing means AlwaysOn."""
    from opentelemetry.sdk.trace import sampling as s

    if kind is None:
        return s.ParentBased(s.TraceIdRatioBased(ratio)) if ratio is not None else s.ALWAYS_ON
    k = kind.strip().lower()
    if k not in SAMPLERS:
        raise OtlpConfigError(f"unknown sampler {kind!r}")
    if k == "always_on":
        return s.ALWAYS_ON
    if k == "always_off":
        return s.ALWAYS_OFF
    r = 1.0 if ratio is None else float(ratio)
    return s.TraceIdRatioBased(r) if k == "traceidratio" else s.ParentBased(s.TraceIdRatioBased(r))


class JsonlSpanExporter:
    """Structured JSONL span sink (one line per finished span)."""

    def __init__(self, path: str):
        self._f = open(path, "a", buffering=1)
        self._lock = threading.Lock()

    def export(self, spans) -> Any:
        from opentelemetry.sdk.trace.export import SpanExportResult

        with self._lock:
            for sp in spans:
                self._f.write(json.dumps({
                    "name": sp.name, "trace_id": f"{sp.context.trace_id:032x}", "span_id": f"{sp.context.span_id:016x}",
                    "start_ns": sp.start_time, "end_ns": sp.end_time,
                    "attributes": dict(sp.attributes or {})}) + "\n")
        return SpanExportResult.SUCCESS

    def shutdown(self) -> None:
        self._f.close()

    def force_flush(self, timeout_millis: int = 30000) -> bool:
        return True


class Telemetry:
    """Per-process telemetry bundle built from a role config (``telemetry_*`` keys)."""

    def __init__(self, service: str, *, endpoint: Optional[str] = None, protocol: Optional[str] = None,
                 headers: Optional[str] = None, attributes: Optional[str] = None,
                 sampler: Optional[str] = None, sample_ratio: Optional[float] = None,
                 jsonl_path: Optional[str] = None, metric_interval_s: float = 1.0,
                 span_exporter=None, metric_reader=None):
        from opentelemetry.sdk.metrics import MeterProvider
        from opentelemetry.sdk.resources import Resource
        from opentelemetry.sdk.trace import TracerProvider
        from opentelemetry.sdk.trace.export import BatchSpanProcessor, SimpleSpanProcessor

        self.service = service
        self.endpoint = Endpoint.parse(endpoint)
        self.protocol = Protocol.parse(protocol)
        self.headers = parse_kv_list(headers)
        attrs = {"service.name": service}
        attrs.update(parse_kv_list(attributes))
        self.resource = Resource.create(attrs)
        self.enabled = self.endpoint is not None
        self.tracer_provider = TracerProvider(resource=self.resource, sampler=make_sampler(sampler, sample_ratio))
        if span_exporter is not None:
            self.tracer_provider.add_span_processor(SimpleSpanProcessor(span_exporter))
        if jsonl_path:
            self.tracer_provider.add_span_processor(SimpleSpanProcessor(JsonlSpanExporter(jsonl_path)))
        readers = [metric_reader] if metric_reader is not None else []
        if self.enabled:
            self.tracer_provider.add_span_processor(BatchSpanProcessor(self._span_exporter()))
            readers.append(self._metric_reader(metric_interval_s))
        self.meter_provider = MeterProvider(resource=self.resource, metric_readers=readers)
        self.tracer = self.tracer_provider.get_tracer("hypha_b200")
        self.meter = self.meter_provider.get_meter("hypha_b200")
        self._bw_sources: List[BandwidthCounters] = []
        self._bw_lock = threading.Lock()
        self.meter.create_observable_counter(INBOUND, callbacks=[lambda o: self._observe(True)], unit="By")
        self.meter.create_observable_counter(OUTBOUND, callbacks=[lambda o: self._observe(False)], unit="By")
        self.logger = self._make_logger()

    # -- exporters ---------------------------------------------------------------------------
    def _span_exporter(self):
        if self.protocol.value == "grpc":
            from opentelemetry.exporter.otlp.proto.grpc.trace_exporter import OTLPSpanExporter

            return OTLPSpanExporter(endpoint=self.endpoint.url, headers=tuple(self.headers.items()) or None)
        from opentelemetry.exporter.otlp.proto.http.trace_exporter import OTLPSpanExporter

        return OTLPSpanExporter(endpoint=self.endpoint.signal(self.protocol, "traces"), headers=self.headers or None)

    def _metric_reader(self, interval_s: float):
        from opentelemetry.sdk.metrics.export import PeriodicExportingMetricReader

        if self.protocol.value == "grpc":
            from opentelemetry.exporter.otlp.proto.grpc.metric_exporter import OTLPMetricExporter

            exp = OTLPMetricExporter(endpoint=self.endpoint.url, headers=tuple(self.headers.items()) or None)
        else:
            from opentelemetry.exporter.otlp.proto.http.metric_exporter import OTLPMetricExporter

            exp = OTLPMetricExporter(endpoint=self.endpoint.signal(self.protocol, "metrics"), headers=self.headers or None)
        return PeriodicExportingMetricReader(exp, export_interval_millis=int(interval_s * 1000))

    def _make_logger(self) -> logging.Logger:
        log = logging.getLogger(f"hypha.{self.service}")
        if not log.handlers:
            h = logging.StreamHandler(sys.stdout)
            h.setFormatter(logging.Formatter("%(asctime)s %(levelname)s %(name)s %(message)s"))
            log.addHandler(h)
            log.setLevel(os.environ.get("HYPHA_LOG", os.environ.get("RUST_LOG", "INFO")).upper()
                         if os.environ.get("HYPHA_LOG", os.environ.get("RUST_LOG", "INFO")).upper()
                         in ("DEBUG", "INFO", "WARNING", "ERROR") else "INFO")
            log.propagate = False
        if self.enabled:
            try:
                from opentelemetry.sdk._logs import LoggerProvider, LoggingHandler
                from opentelemetry.sdk._logs.export import BatchLogRecordProcessor

                if self.protocol.value == "grpc":
                    from opentelemetry.exporter.otlp.proto.grpc._log_exporter import OTLPLogExporter

                    exp = OTLPLogExporter(endpoint=self.endpoint.url)
                else:
                    from opentelemetry.exporter.otlp.proto.http._log_exporter import OTLPLogExporter

                    exp = OTLPLogExporter(endpoint=self.endpoint.signal(self.protocol, "logs"))
                lp = LoggerProvider(resource=self.resource)
                lp.add_log_record_processor(BatchLogRecordProcessor(exp))
                log.addHandler(LoggingHandler(logger_provider=lp))
                self._log_provider = lp
            except Exception:      # log bridge is best-effort
                pass
        return log

    # -- bandwidth ------------------------------------------------------------------------------
    def register_bandwidth(self, counters: BandwidthCounters) -> None:
        with self._bw_lock:
            self._bw_sources.append(counters)

    def _observe(self, inbound: bool):
        from opentelemetry.metrics import Observation

        agg: Dict[str, int] = {}
        with self._bw_lock:
            for c in self._bw_sources:
                snap = c.snapshot()[INBOUND if inbound else OUTBOUND]
                for proto, n in snap.items():
                    agg[proto] = agg.get(proto, 0) + n
        return [Observation(n, {"protocols": proto}) for proto, n in agg.items()]

    # -- spans -------------------------------------------------------------------------------------
    @contextlib.contextmanager
    def span(self, name: str, **attrs) -> Iterator[Any]:
        with self.tracer.start_as_current_span(name, attributes=attrs) as sp:
            yield sp

    def shutdown(self) -> None:
        for p in (self.tracer_provider, self.meter_provider, getattr(self, "_log_provider", None)):
            try:
                if p is not None:
                    p.shutdown()
            except Exception:
                pass

    @classmethod
    def from_config(cls, service: str, cfg: Any, **kw) -> "Telemetry":
        return cls(service, endpoint=getattr(cfg, "telemetry_endpoint", None),
                   protocol=getattr(cfg, "telemetry_protocol", None), headers=getattr(cfg, "telemetry_headers", None),
                   attributes=getattr(cfg, "telemetry_attributes", None), sampler=getattr(cfg, "telemetry_sampler", None),
                   sample_ratio=getattr(cfg, "telemetry_sample_ratio", None), **kw)


# -- device-side timing ----------------------------------------------------------------------------
@contextlib.contextmanager
def nvtx_range(name: str) -> Iterator[None]:
    """NVTX range when CUDA is present (shows up in ncu/nsys timelines), else a no-op."""
    try:
        import torch

        on = torch.cuda.is_available()
    except Exception:
        on = False
    if on:
        torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        if on:
            torch.cuda.nvtx.range_pop()


class DeviceTimer:
    """CUDA-event stopwatch on the current stream: ``with timer: ...`` then ``timer.ms``."""

    def __init__(self) -> None:
        import torch

        self._torch = torch
        self._e0 = torch.cuda.Event(enable_timing=True)
        self._e1 = torch.cuda.Event(enable_timing=True)
        self.ms: Optional[float] = None

    def __enter__(self) -> "DeviceTimer":
        self._e0.record()
        return self

    def __exit__(self, *exc) -> None:
        self._e1.record()
        self._e1.synchronize()
        self.ms = self._e0.elapsed_time(self._e1)


def max_over_ranks(value: float, group=None) -> float:
    """Reduce a per-rank device time to the job-level number (max over ranks)."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return value
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


__all__ = ["BandwidthCounters", "DeviceTimer", "Endpoint", "JsonlSpanExporter", "OtlpConfigError", "Protocol",
           "Telemetry", "format_kv_list", "make_sampler", "max_over_ranks", "nvtx_range", "parse_kv_list",
           "INBOUND", "OUTBOUND", "SAMPLERS"]
