"""Layered role configuration: defaults -> TOML file -> ``HYPHA_*`` env -> ``OTEL_*`` env -> CLI.

Reference: crates/config/src/lib.rs (builder :565-616, metadata / provenance :403-450, validation
hook :453-462, TLS loaders :464-541, ``to_toml`` with doc comments :544-563) and the per-role
``config.rs`` files (key sets: SURVEY appendix C).  Precedence, key names, env prefixes and the
OTEL aliases (scheduler/src/config.rs:117-178) are the same; diagnostics point at the TOML line
that supplied an offending value like the reference's miette spans.
"""
from __future__ import annotations

import ipaddress
import json
import os
import tomllib
from dataclasses import MISSING, dataclass, field, fields, is_dataclass
from typing import Any, Dict, List, Optional, Tuple, Type, TypeVar, get_args, get_origin, get_type_hints

T = TypeVar("T")

# reserved ranges excluded from the DHT routing table by default (crates/network/src/lib.rs:57-98)
DEFAULT_EXCLUDE_CIDR = [
    "127.0.0.0/8", "10.0.0.0/8", "172.16.0.0/12", "192.168.0.0/16", "169.254.0.0/16",
    "::1/128", "fc00::/7", "fe80::/10",
]

OTEL_ALIASES = {
    "exporter_otlp_endpoint": "telemetry_endpoint",
    "exporter_otlp_headers": "telemetry_headers",
    "exporter_otlp_protocol": "telemetry_protocol",
    "resource_attributes": "telemetry_attributes",
    "traces_sampler": "telemetry_sampler",
    "traces_sampler_arg": "telemetry_sample_ratio",
}


class ConfigError(Exception):
    """Configuration failure with provenance (file, line, key) when known."""

    def __init__(self, message: str, *, key: Optional[str] = None, source: Optional[str] = None,
                 line: Optional[int] = None, snippet: Optional[str] = None):
        self.key, self.source, self.line, self.snippet = key, source, line, snippet
        where = ""
        if source:
            where = f" [{source}" + (f":{line}" if line else "") + "]"
        body = f"{message}{where}"
        if snippet:
            body += f"\n  {line or ''} | {snippet}"
        super().__init__(body)


def doc(text: str, default: Any = MISSING, *, factory=None, alias: Optional[str] = None):
    md = {"doc": text}
    if alias:
        md["alias"] = alias
    if factory is not None:
        return field(default_factory=factory, metadata=md)
    return field(default=default, metadata=md)


# ---------------------------------------------------------------------------------------------
# schema dataclasses
# ---------------------------------------------------------------------------------------------
@dataclass
class TelemetryMixin:
    telemetry_endpoint: Optional[str] = doc("OTLP endpoint (OTEL_EXPORTER_OTLP_ENDPOINT); unset disables export", None)
    telemetry_attributes: Optional[str] = doc("resource attributes `k=v,k=v` (OTEL_RESOURCE_ATTRIBUTES)", None)
    telemetry_headers: Optional[str] = doc("exporter headers `k=v,k=v` (OTEL_EXPORTER_OTLP_HEADERS)", None)
    telemetry_protocol: Optional[str] = doc("`grpc` or `http/protobuf` (OTEL_EXPORTER_OTLP_PROTOCOL)", None)
    telemetry_sampler: Optional[str] = doc("always_on | always_off | traceidratio | parentbased_traceidratio", None)
    telemetry_sample_ratio: Optional[float] = doc("sampling ratio for the ratio samplers (OTEL_TRACES_SAMPLER_ARG)", None)


@dataclass
class NetworkMixin(TelemetryMixin):
    cert_pem: str = doc("path to this node's certificate chain (PEM)", "")
    key_pem: str = doc("path to this node's private key (PEM, PKCS#8 Ed25519)", "")
    trust_pem: str = doc("path to the trusted CA bundle (PEM)", "")
    crls_pem: Optional[str] = doc("optional certificate revocation lists (PEM)", None)
    listen_addresses: List[str] = doc("addresses to listen on (memory://, unix://, tcp://host:port, udp://host:port)",
                                      factory=lambda: ["tcp://127.0.0.1:0"])
    external_addresses: List[str] = doc("addresses advertised to peers", factory=list)
    exclude_cidr: List[str] = doc("CIDR ranges never added to the routing table",
                                  factory=lambda: list(DEFAULT_EXCLUDE_CIDR))


@dataclass
class GatewayConfig(NetworkMixin):
    """hypha-gateway: bootstrap / relay / DHT-server node."""


@dataclass
class ResourcesSection:
    gpu: float = doc("GPU units offered", 16.0)
    cpu: float = doc("CPU cores offered", 1.0)
    memory: float = doc("memory offered (GB)", 8.0)
    storage: float = doc("scratch storage offered (GB)", 20.0)


@dataclass
class PriceRange:
    bid: float = doc("opening bid per weighted resource unit", 100.0)
    max: float = doc("highest acceptable counter-offer", 100.0)


@dataclass
class OfferStrategy:
    type: str = doc("`whole` (offer everything at max(price, bid)) or `flexible` (mirror the request)", "flexible")


@dataclass
class OfferSection:
    price: float = doc("asking price", 100.0)
    floor: float = doc("bids below this are ignored", 100.0)
    strategy: OfferStrategy = doc("offer strategy", factory=OfferStrategy)


@dataclass
class ExecutorSection:
    cls: str = doc("executor class: `train` or `aggregate`", "train", alias="class")
    name: str = doc("executor name matched against the job's descriptor", "diloco-transformer")
    runtime: str = doc("`process` (spawn cmd) or `parameter-server` (built in)", "process")
    cmd: Optional[str] = doc("command for the process runtime", None)
    args: List[str] = doc("arguments; {SOCKET_PATH} {WORK_DIR} {JOB_JSON} are substituted", factory=list)
    env: Dict[str, str] = doc("extra environment of the spawned process (e.g. CUDA_VISIBLE_DEVICES)", factory=dict)


def _default_executors() -> List[ExecutorSection]:
    return [
        ExecutorSection("train", "diloco-transformer", "process", "python",
                        ["-m", "hypha_b200.executor.training", "--socket", "{SOCKET_PATH}",
                         "--work-dir", "{WORK_DIR}", "--job", "{JOB_JSON}"]),
        ExecutorSection("aggregate", "parameter-server", "parameter-server", None, []),
    ]


@dataclass
class WorkerConfig(NetworkMixin):
    """hypha-worker: resource owner + job runner."""

    gateway_addresses: List[str] = doc("gateways to dial at start-up", factory=list)
    relay_circuit: bool = doc("listen on a relay circuit through each gateway", True)
    work_dir: str = doc("base directory for per-job work dirs", "/tmp")
    device: Optional[int] = doc("CUDA device index this worker's executors train on (exported to them as "
                                "HYPHA_DEVICE; one worker per GPU on a multi-GPU box)", None)
    resources: ResourcesSection = doc("capacity offered to schedulers", factory=ResourcesSection)
    offer: OfferSection = doc("pricing", factory=OfferSection)
    executors: List[ExecutorSection] = doc("executors this worker can run", factory=_default_executors)


@dataclass
class DataConfig(NetworkMixin):
    """hypha-data: announces a dataset directory and serves its slices."""

    gateway_addresses: List[str] = doc("gateways to dial at start-up", factory=list)
    dataset_path: str = doc("directory whose files are the slices; its name is the dataset name", "")


@dataclass
class ModelSource:
    repository: str = doc("model repository", "hypha-space/lenet")
    revision: Optional[str] = doc("revision", None)
    filenames: List[str] = doc("files to fetch", factory=lambda: ["config.json", "model.safetensors",
                                                                  "configuration_lenet.py", "modeling_lenet.py"])
    token: Optional[str] = doc("access token", None)
    type: str = doc("model task (ModelType, kebab-case)", "image-classification")
    input_names: List[str] = doc("model input names", factory=lambda: ["pixel_values", "labels"])


@dataclass
class PreprocessorSource:
    repository: str = doc("preprocessor repository (empty = the job has no preprocessor)", "hypha-space/lenet")
    revision: Optional[str] = doc("revision", None)
    filenames: List[str] = doc("files to fetch", factory=lambda: ["preprocessor_config.json"])
    token: Optional[str] = doc("access token", None)
    type: str = doc("tokenizer | feature | image | video | auto", "image")
    input_names: List[str] = doc("preprocessor input names", factory=lambda: ["images"])


@dataclass
class DatasetSource:
    dataset: str = doc("dataset name announced by a data node", "mnist")


@dataclass
class Rounds:
    avg_samples_between_updates: int = doc("samples processed by all workers between outer steps", 1200)
    update_rounds: int = doc("number of DiLoCo outer rounds", 100)
    max_batch_size: Optional[int] = doc("cap on the per-worker batch size", 600)


@dataclass
class AdamSection:
    learning_rate: float = doc("inner AdamW learning rate", 1e-3, alias="learning-rate")
    betas: Optional[List[float]] = doc("AdamW betas", None)
    epsilon: Optional[float] = doc("AdamW epsilon", None)


@dataclass
class NesterovSection:
    learning_rate: float = doc("outer learning rate", 0.7, alias="learning-rate")
    momentum: float = doc("outer momentum", 0.9)


def _worker_res() -> ResourcesSection:
    return ResourcesSection(gpu=10.0, cpu=1.0, memory=1.0, storage=0.0)


def _ps_res() -> ResourcesSection:
    return ResourcesSection(gpu=0.0, cpu=1.0, memory=1.0, storage=0.0)


@dataclass
class JobResources:
    num_workers: int = doc("training workers to allocate", 2)
    worker: ResourcesSection = doc("minimum resources per training worker", factory=_worker_res)
    parameter_server: ResourcesSection = doc("resources for the aggregate role", factory=_ps_res)
    worker_price: PriceRange = doc("price range for training workers", factory=PriceRange)
    parameter_server_price: PriceRange = doc("price range for the aggregate role", factory=PriceRange)


@dataclass
class CheckpointSection:
    dir: Optional[str] = doc("directory (on the aggregate worker) where the job's outer state is saved: momentum "
                             "and the cumulative update U, from which any worker rebuilds theta = theta_0(seed) + U", None)
    every: int = doc("save after every N-th round", 1)
    resume: bool = doc("start from the newest checkpoint in `dir` instead of round 0", False)


@dataclass
class ElasticSection:
    enabled: bool = doc("keep the job running when a train worker's lease is lost (the reference scheduler "
                        "exits on the first failure)", True)
    min_workers: int = doc("stop the job when fewer train workers than this remain", 1)
    replace: bool = doc("try to lease a replacement for a lost train worker and let it join the running job "
                        "(parameter-server mode)", True)
    max_replacements: int = doc("stop replacing after this many replacements (a job whose executor always "
                                "crashes must not be restarted forever)", 4)


@dataclass
class AggregationSection:
    mode: str = doc("`parameter-server` (reference behaviour: pseudo-gradient files through an aggregate worker) "
                    "or `collective` (the train workers synchronise among themselves: fused NVLink outer step on "
                    "GPUs, gloo all-reduce on CPU; no aggregate worker is allocated)", "parameter-server")
    rendezvous: Optional[str] = doc("`tcp://host:port` the ranks of a collective job meet on "
                                    "(default: a free port on 127.0.0.1)", None)
    weighting: str = doc("`uniform` (every worker's pseudo-gradient counts 1/N, like the reference's intent) or "
                         "`samples` (weighted by the samples each worker processed in the round; matters when "
                         "performance-aware pacing gives workers different shares)", "uniform")


@dataclass
class PrecisionSection:
    gemm: str = doc("precision of the Linear-layer GEMMs of the inner steps on GPU executors: `bf16` (default), "
                    "`mxfp8` (e4m3 with one UE8M0 scale per 32 elements, block-scaled tensor-core MMA) or `fp8` "
                    "(per-tensor scaling); master weights, optimizer and the outer step are unaffected", "bf16")
    micro_batches: int = doc("gradient accumulation: every batch of an inner step is split into this many "
                             "micro-batches whose mean gradient feeds ONE optimizer step (1 = off, the reference's "
                             "behaviour); the worker's batch size must be a multiple of it", 1)
    main_grad: str = doc("where the micro-batch gradients are summed when `micro_batches` > 1: `fp32` (an fp32 "
                         "buffer, rounded to bf16 once; +4 bytes per parameter) or `bf16` (in the bf16 gradient "
                         "buffer by the wgrad GEMM epilogues; no extra memory, one rounding per micro-batch)", "fp32")


@dataclass
class DiLoCoJob:
    type: str = doc("job type (only `diloco`)", "diloco")
    model: ModelSource = doc("model source", factory=ModelSource)
    preprocessor: Optional[PreprocessorSource] = doc("optional preprocessor", factory=PreprocessorSource)
    dataset: DatasetSource = doc("dataset", factory=DatasetSource)
    rounds: Rounds = doc("round structure", factory=Rounds)
    inner_optimizer: AdamSection = doc("inner optimizer", factory=AdamSection)
    outer_optimizer: NesterovSection = doc("outer optimizer", factory=NesterovSection)
    resources: JobResources = doc("allocation request", factory=JobResources)
    aggregation: AggregationSection = doc("how the outer step is carried out", factory=AggregationSection)
    elastic: ElasticSection = doc("behaviour on worker failure", factory=ElasticSection)
    checkpoint: CheckpointSection = doc("job-level checkpoint / resume (parameter-server mode)", factory=CheckpointSection)
    precision: PrecisionSection = doc("numeric precision of the inner steps", factory=PrecisionSection)


@dataclass
class SchedulerSection:
    job: DiLoCoJob = doc("the job this scheduler owns", factory=DiLoCoJob)


@dataclass
class SchedulerConfig(NetworkMixin):
    """hypha-scheduler: job owner."""

    gateway_addresses: List[str] = doc("gateways to dial at start-up", factory=list)
    relay_circuit: bool = doc("listen on a relay circuit through each gateway", True)
    status_bridge: Optional[str] = doc("`host:port` of the AIM driver receiving metrics", None)
    scheduler: SchedulerSection = doc("scheduler section", factory=SchedulerSection)


# ---------------------------------------------------------------------------------------------
# dataclass <-> dict
# ---------------------------------------------------------------------------------------------
def _key(f) -> str:
    return f.metadata.get("alias", f.name)


def to_dict(obj: Any) -> Any:
    if is_dataclass(obj):
        return {_key(f): to_dict(getattr(obj, f.name)) for f in fields(obj)}
    if isinstance(obj, list):
        return [to_dict(x) for x in obj]
    return obj


def _strip_optional(tp):
    if get_origin(tp) is Optional or (get_origin(tp) is not None and type(None) in get_args(tp)):
        args = [a for a in get_args(tp) if a is not type(None)]
        return args[0] if len(args) == 1 else tp
    return tp


def from_dict(cls: Type[T], data: Dict[str, Any], path: str = "", meta: "Optional[Provenance]" = None) -> T:
    hints = get_type_hints(cls)
    kwargs = {}
    known = {}
    for f in fields(cls):
        known[_key(f)] = f
        if "alias" in f.metadata:
            known.setdefault(f.name, f)          # accept snake_case too
    for k, v in data.items():
        if k not in known:
            raise _err(meta, f"{path}{k}", f"unknown configuration key `{path}{k}`")
        f = known[k]
        tp = _strip_optional(hints[f.name])
        kwargs[f.name] = _coerce(tp, v, f"{path}{k}", meta)
    try:
        return cls(**kwargs)
    except TypeError as exc:
        raise ConfigError(str(exc)) from exc


def _coerce(tp, v, path, meta):
    if v is None:
        return None
    origin = get_origin(tp)
    try:
        if is_dataclass(tp):
            if not isinstance(v, dict):
                raise ValueError("expected a table")
            return from_dict(tp, v, path + ".", meta)
        if origin in (list, List):
            (inner,) = get_args(tp) or (Any,)
            if isinstance(v, str):
                v = [x for x in (s.strip() for s in v.strip("[]").split(",")) if x]
            return [_coerce(_strip_optional(inner), x, f"{path}[{i}]", meta) for i, x in enumerate(v)]
        if tp is bool:
            if isinstance(v, str):
                return v.strip().lower() in ("1", "true", "yes", "on")
            return bool(v)
        if tp is int:
            if isinstance(v, bool) or (isinstance(v, float) and v != int(v)):
                raise ValueError("expected an integer")
            return int(v)
        if tp is float:
            return float(v)
        if tp is str:
            if not isinstance(v, str):
                raise ValueError("expected a string")
            return v
    except ConfigError:
        raise
    except (ValueError, TypeError) as exc:
        raise _err(meta, path, f"invalid value for `{path}`: {exc}") from exc
    return v


# ---------------------------------------------------------------------------------------------
# provenance + layering
# ---------------------------------------------------------------------------------------------
class Provenance:
    """Remembers which layer supplied each key (and the TOML line for diagnostics)."""

    def __init__(self) -> None:
        self.source: Dict[str, str] = {}
        self.toml_path: Optional[str] = None
        self.toml_lines: List[str] = []

    def set(self, key: str, source: str) -> None:
        self.source[key] = source

    def locate(self, key: str) -> Tuple[Optional[str], Optional[int], Optional[str]]:
        src = self.source.get(key) or self.source.get(key.split(".")[0].split("[")[0])
        if src and src.startswith("file:") and self.toml_lines:
            leaf = key.split(".")[-1].split("[")[0]
            for i, line in enumerate(self.toml_lines, 1):
                s = line.split("#", 1)[0].strip()
                if s.startswith(leaf) and "=" in s and s.split("=", 1)[0].strip() == leaf:
                    return self.toml_path, i, line.rstrip()
            return self.toml_path, None, None
        return src, None, None


def _err(meta: Optional[Provenance], key: str, message: str) -> ConfigError:
    if meta is None:
        return ConfigError(message, key=key)
    source, line, snippet = meta.locate(key)
    return ConfigError(message, key=key, source=source, line=line, snippet=snippet)


def _deep_merge(base: dict, over: dict) -> dict:
    out = dict(base)
    for k, v in over.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = _deep_merge(out[k], v)
        else:
            out[k] = v
    return out


def _flatten_keys(d: dict, prefix: str = "") -> List[str]:
    out = []
    for k, v in d.items():
        out.append(prefix + k)
        if isinstance(v, dict):
            out += _flatten_keys(v, prefix + k + ".")
    return out


class ConfigBuilder:
    """``builder(WorkerConfig).with_toml(path).with_env().with_overrides(cli).build()``"""

    def __init__(self, cls: Type[T]):
        self.cls = cls
        self.meta = Provenance()
        self.data: dict = to_dict(cls())
        for k in _flatten_keys(self.data):
            self.meta.set(k, "default")

    def _layer(self, over: dict, source: str) -> "ConfigBuilder":
        self.data = _deep_merge(self.data, over)
        for k in _flatten_keys(over):
            self.meta.set(k, source)
        return self

    def with_toml(self, path: Optional[str], required: bool = True) -> "ConfigBuilder":
        if not path:
            return self
        if not os.path.exists(path):
            if required:
                raise ConfigError(f"configuration file not found: {path}", source=path)
            return self
        text = open(path, "r", encoding="utf-8").read()
        try:
            parsed = tomllib.loads(text)
        except tomllib.TOMLDecodeError as exc:
            raise ConfigError(f"invalid TOML: {exc}", source=f"file:{path}") from exc
        self.meta.toml_path = path
        self.meta.toml_lines = text.splitlines()
        return self._layer(parsed, f"file:{path}")

    def with_env(self, environ: Optional[Dict[str, str]] = None) -> "ConfigBuilder":
        env = os.environ if environ is None else environ
        over: dict = {}
        top = {_key(f): f for f in fields(self.cls)}
        for name, value in sorted(env.items()):
            if name.startswith("HYPHA_"):
                # HYPHA_<KEY> for top-level keys (the reference's figment Env::prefixed); nested keys are
                # reachable either as JSON (HYPHA_SCHEDULER='{"job": ...}') or with `__` as the separator
                # (HYPHA_SCHEDULER__JOB__ROUNDS__UPDATE_ROUNDS=5)
                path = name[len("HYPHA_"):].lower().split("__")
                if path[0] not in top:
                    continue
                node = over
                for part in path[:-1]:
                    nxt = node.get(part)
                    if not isinstance(nxt, dict):
                        nxt = node[part] = {}
                    node = nxt
                val = _parse_env_value(value)
                if len(path) > 1 and isinstance(val, str):
                    val = _parse_scalar(val)
                node[path[-1]] = val
        self._layer(over, "env:HYPHA_*")
        over = {}
        for name, value in env.items():
            if name.startswith("OTEL_"):
                key = OTEL_ALIASES.get(name[len("OTEL_"):].lower())
                if key and key in top:
                    over[key] = value
        return self._layer(over, "env:OTEL_*")

    def with_overrides(self, values: Optional[Dict[str, Any]]) -> "ConfigBuilder":
        if not values:
            return self
        clean = {k: v for k, v in values.items() if v is not None}
        return self._layer(clean, "cli")

    def build(self) -> T:
        cfg = from_dict(self.cls, self.data, meta=self.meta)
        cfg._meta = self.meta  # type: ignore[attr-defined]
        # TOML has no null: an empty `repository` (or `enabled = false` style `type = "none"`) under
        # [scheduler.job.preprocessor] means "this job has no preprocessor"
        job = getattr(getattr(cfg, "scheduler", None), "job", None)
        pre = getattr(job, "preprocessor", None)
        if pre is not None and (not pre.repository or pre.type == "none"):
            job.preprocessor = None
        return cfg


def _parse_scalar(v: str) -> Any:
    """Nested `__` variables carry scalars: numbers and booleans are recognised, the rest stays text."""
    low = v.strip().lower()
    if low in ("true", "false"):
        return low == "true"
    for cast in (int, float):
        try:
            return cast(v)
        except ValueError:
            pass
    return v


def _parse_env_value(v: str) -> Any:
    s = v.strip()
    if s.startswith("[") or s.startswith("{"):
        try:
            return json.loads(s)
        except json.JSONDecodeError:
            pass
    return v


def builder(cls: Type[T]) -> ConfigBuilder:
    return ConfigBuilder(cls)


# ---------------------------------------------------------------------------------------------
# validation
# ---------------------------------------------------------------------------------------------
def _addr_ip(addr: str) -> Optional[str]:
    if addr.startswith("tcp://"):
        host = addr[len("tcp://"):].rsplit(":", 1)[0]
        return host.strip("[]")
    if addr.startswith("/ip4/") or addr.startswith("/ip6/"):
        return addr.split("/")[2]
    return None


def validate(cfg: Any) -> Any:
    """Role-independent checks; raises ``ConfigError`` with the TOML location when known.

    * gateway addresses must not fall into ``exclude_cidr`` (worker/src/config.rs:578-595);
    * executor sections must be well-formed; offer strategy / model types must be known."""
    meta: Optional[Provenance] = getattr(cfg, "_meta", None)
    nets = []
    for c in getattr(cfg, "exclude_cidr", []) or []:
        try:
            nets.append(ipaddress.ip_network(c, strict=False))
        except ValueError as exc:
            raise _err(meta, "exclude_cidr", f"invalid CIDR `{c}`: {exc}") from exc
    for addr in getattr(cfg, "gateway_addresses", []) or []:
        ip = _addr_ip(addr)
        if ip is None:
            continue
        try:
            ipa = ipaddress.ip_address(ip)
        except ValueError:
            continue            # host names are resolved at dial time
        for n in nets:
            if ipa.version == n.version and ipa in n:
                raise _err(meta, "gateway_addresses",
                           f"gateway address `{addr}` is inside excluded range `{n}`; remove the range "
                           f"from `exclude_cidr` or use a routable address")
    if isinstance(cfg, WorkerConfig):
        if cfg.offer.strategy.type not in ("whole", "flexible"):
            raise _err(meta, "offer.strategy.type", f"unknown offer strategy `{cfg.offer.strategy.type}`")
        for i, e in enumerate(cfg.executors):
            if e.cls not in ("train", "aggregate"):
                raise _err(meta, f"executors[{i}].class", f"unknown executor class `{e.cls}`")
            if e.runtime not in ("process", "parameter-server"):
                raise _err(meta, f"executors[{i}].runtime", f"unknown runtime `{e.runtime}`")
            if e.runtime == "process" and not e.cmd:
                raise _err(meta, f"executors[{i}].cmd", "process executors need `cmd`")
    if isinstance(cfg, SchedulerConfig):
        from .messages import MODEL_TYPES, PREPROCESSOR_TYPES

        job = cfg.scheduler.job
        if job.type != "diloco":
            raise _err(meta, "scheduler.job.type", f"unknown job type `{job.type}`")
        if job.model.type not in MODEL_TYPES:
            raise _err(meta, "scheduler.job.model.type", f"unknown model type `{job.model.type}`")
        if job.preprocessor is not None and job.preprocessor.type not in PREPROCESSOR_TYPES:
            raise _err(meta, "scheduler.job.preprocessor.type", f"unknown preprocessor type `{job.preprocessor.type}`")
        if job.resources.num_workers < 1:
            raise _err(meta, "scheduler.job.resources.num_workers", "need at least one worker")
        if job.aggregation.mode not in ("parameter-server", "collective"):
            raise _err(meta, "scheduler.job.aggregation.mode",
                       f"unknown aggregation mode `{job.aggregation.mode}` (parameter-server | collective)")
        if job.precision.gemm not in ("bf16", "mxfp8", "fp8"):
            raise _err(meta, "scheduler.job.precision.gemm",
                       f"unknown GEMM precision `{job.precision.gemm}` (bf16 | mxfp8 | fp8)")
        if not isinstance(job.precision.micro_batches, int) or job.precision.micro_batches < 1:
            raise _err(meta, "scheduler.job.precision.micro_batches", "must be a positive integer")
        if job.precision.main_grad not in ("fp32", "bf16"):
            raise _err(meta, "scheduler.job.precision.main_grad",
                       f"unknown main-gradient type `{job.precision.main_grad}` (fp32 | bf16)")
        if job.aggregation.weighting not in ("uniform", "samples"):
            raise _err(meta, "scheduler.job.aggregation.weighting",
                       f"unknown weighting `{job.aggregation.weighting}` (uniform | samples)")
        if job.aggregation.rendezvous and not job.aggregation.rendezvous.startswith("tcp://"):
            raise _err(meta, "scheduler.job.aggregation.rendezvous", "rendezvous must be a tcp://host:port address")
    if isinstance(cfg, DataConfig) and cfg.dataset_path and not os.path.isdir(cfg.dataset_path):
        raise _err(meta, "dataset_path", f"dataset_path `{cfg.dataset_path}` is not a directory")
    return cfg


# ---------------------------------------------------------------------------------------------
# documented TOML emission (`init`)
# ---------------------------------------------------------------------------------------------
def _toml_value(v: Any) -> str:
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, (int, float)):
        return repr(v)
    if isinstance(v, str):
        return json.dumps(v)
    if isinstance(v, list):
        return "[" + ", ".join(_toml_value(x) for x in v) + "]"
    if isinstance(v, dict):                      # inline table, e.g. an executor's `env`
        return "{ " + ", ".join(f"{json.dumps(str(k))} = {_toml_value(x)}" for k, x in v.items()) + " }" if v else "{}"
    raise TypeError(type(v))


def to_toml(cfg: Any) -> str:
    """Render a config as TOML with each field's documentation as a ``#`` comment."""
    lines: List[str] = []
    cls_doc = (type(cfg).__doc__ or "").strip()
    if cls_doc:
        lines += [f"# {l.strip()}" for l in cls_doc.splitlines()] + [""]

    def emit(obj, prefix: str) -> None:
        scalars, tables, arrays = [], [], []
        for f in fields(obj):
            v = getattr(obj, f.name)
            if is_dataclass(v):
                tables.append((f, v))
            elif isinstance(v, list) and v and is_dataclass(v[0]):
                arrays.append((f, v))
            else:
                scalars.append((f, v))
        for f, v in scalars:
            d = f.metadata.get("doc")
            if d:
                lines.append(f"# {d}")
            if v is None:
                lines.append(f"# {_key(f)} = ...")
            else:
                lines.append(f"{_key(f)} = {_toml_value(v)}")
        for f, v in tables:
            lines.append("")
            d = f.metadata.get("doc")
            if d:
                lines.append(f"# {d}")
            lines.append(f"[{prefix}{_key(f)}]")
            emit(v, f"{prefix}{_key(f)}.")
        for f, vs in arrays:
            for v in vs:
                lines.append("")
                d = f.metadata.get("doc")
                if d:
                    lines.append(f"# {d}")
                lines.append(f"[[{prefix}{_key(f)}]]")
                emit(v, f"{prefix}{_key(f)}.")

    emit(cfg, "")
    return "\n".join(lines) + "\n"


# ---------------------------------------------------------------------------------------------
# TLS material (accepted for API parity; a single-box loopback fabric does not need mTLS)
# ---------------------------------------------------------------------------------------------
def load_tls(cfg: NetworkMixin) -> Dict[str, Optional[bytes]]:
    """Read cert/key/trust/CRL PEM files when configured (crates/config/src/lib.rs:464-541)."""
    out: Dict[str, Optional[bytes]] = {}
    for name in ("cert_pem", "key_pem", "trust_pem", "crls_pem"):
        path = getattr(cfg, name)
        if path:
            try:
                out[name] = open(path, "rb").read()
            except OSError as exc:
                raise _err(getattr(cfg, "_meta", None), name, f"cannot read `{name}` at {path}: {exc}") from exc
        else:
            out[name] = None
    return out
