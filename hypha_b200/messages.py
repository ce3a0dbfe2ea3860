"""Wire + job data model (JSON / CBOR), shape-compatible with the reference.

Reference: crates/messages/src/lib.rs
  api (:15-44), health (:47-63), progress (:66-119), request_worker (:122-134),
  worker_offer (:137-157), renew_lease (:160-179), dispatch_job (:181-201),
  job_status (:203-214), JobSpec/WorkerSpec (:216-230), Reference + Fetch/Send/Receive
  (:241-417), ModelType (:419-460), executor configs (:490-515), Executor enums (:573-643),
  optimizers / LR schedules (:645-687), parameter_pull/push/data (:698-757), records (:759-775).

Design: plain dataclasses with ``to_wire()`` / ``from_wire()`` producing exactly the
serde shapes (externally-tagged enums for the RPC envelopes, ``type``/``class``-tagged
enums where the reference says so, kebab-case field names where it says so).  Peers are
strings, ids are UUID strings, times are float seconds (wire: serde ``SystemTime``).
"""
from __future__ import annotations

import io
import json
import uuid
from dataclasses import dataclass, field
from enum import Enum
from typing import Any, Dict, List, Optional, Tuple, Union

import cbor2

API_PROTOCOL = "/hypha-api/0.0.1"
HEALTH_PROTOCOL = "/hypha-health/0.0.1"
PROGRESS_PROTOCOL = "/hypha-progress/0.0.1"
IDENTIFY_PROTOCOL = "/hypha-identify/0.0.1"
PUSH_STREAM_PROTOCOL = "/hypha-tensor-stream/push"
PULL_STREAM_PROTOCOL = "/hypha-tensor-stream/pull"
WORKER_TOPIC = "hypha/worker"

TRAIN_EXECUTOR_NAME = "diloco-transformer"      # crates/scheduler/src/bin/hypha-scheduler.rs:47-48
AGGREGATE_EXECUTOR_NAME = "parameter-server"


class MessageError(ValueError):
    pass


def new_id() -> str:
    return str(uuid.uuid4())


# -- time ---------------------------------------------------------------------------------
def time_to_wire(t: float) -> Dict[str, int]:
    secs = int(t)
    return {"secs_since_epoch": secs, "nanos_since_epoch": int(round((t - secs) * 1e9))}


def time_from_wire(d: Any) -> float:
    if isinstance(d, (int, float)):
        return float(d)
    return float(d["secs_since_epoch"]) + float(d["nanos_since_epoch"]) * 1e-9


# -- resources ----------------------------------------------------------------------------
@dataclass
class Resources:
    gpu: float = 0.0
    cpu: float = 0.0
    storage: float = 0.0
    memory: float = 0.0

    def to_wire(self) -> dict:
        return {"gpu": self.gpu, "cpu": self.cpu, "storage": self.storage, "memory": self.memory}

    @classmethod
    def from_wire(cls, d: dict) -> "Resources":
        return cls(float(d.get("gpu", 0)), float(d.get("cpu", 0)), float(d.get("storage", 0)),
                   float(d.get("memory", 0)))

    def to_core(self):
        from .ops.native import load_core

        return load_core().Resources(gpu=self.gpu, cpu=self.cpu, storage=self.storage, memory=self.memory)

    @classmethod
    def from_core(cls, r) -> "Resources":
        return cls(r.gpu, r.cpu, r.storage, r.memory)


# -- references ---------------------------------------------------------------------------
class SelectionStrategy(str, Enum):
    ALL = "All"
    RANDOM = "Random"
    ONE = "One"


@dataclass
class DataSlice:
    dataset: str
    index: int

    def to_wire(self) -> dict:
        return {"dataset": self.dataset, "index": self.index}

    @classmethod
    def from_wire(cls, d: dict) -> "DataSlice":
        return cls(d["dataset"], int(d["index"]))


@dataclass
class DataRecord:
    num_slices: int

    def to_wire(self) -> dict:
        return {"num_slices": self.num_slices}

    @classmethod
    def from_wire(cls, d: dict) -> "DataRecord":
        return cls(int(d["num_slices"]))


@dataclass
class Reference:
    """``type``-tagged: uri | huggingface | peers | scheduler."""

    type: str
    value: Optional[str] = None                       # uri
    repository: Optional[str] = None                  # huggingface
    revision: Optional[str] = None
    filenames: List[str] = field(default_factory=list)
    token: Optional[str] = None
    peers: List[str] = field(default_factory=list)    # peers
    strategy: Optional[SelectionStrategy] = None
    resource: Optional[DataSlice] = None
    peer: Optional[str] = None                        # scheduler
    dataset: Optional[str] = None

    def to_wire(self) -> dict:
        if self.type == "uri":
            return {"type": "uri", "value": self.value}
        if self.type == "huggingface":
            return {"type": "huggingface", "repository": self.repository, "revision": self.revision,
                    "filenames": list(self.filenames), "token": self.token}
        if self.type == "peers":
            return {"type": "peers", "peers": list(self.peers),
                    "strategy": {"type": (self.strategy or SelectionStrategy.ALL).value},
                    "resource": self.resource.to_wire() if self.resource else None}
        if self.type == "scheduler":
            return {"type": "scheduler", "peer": self.peer, "dataset": self.dataset}
        raise MessageError(f"unknown reference type {self.type!r}")

    @classmethod
    def from_wire(cls, d: dict) -> "Reference":
        t = d.get("type")
        if t == "uri":
            return cls("uri", value=d["value"])
        if t == "huggingface":
            return cls("huggingface", repository=d["repository"], revision=d.get("revision"),
                       filenames=list(d.get("filenames", [])), token=d.get("token"))
        if t == "peers":
            strat = d.get("strategy") or {"type": "All"}
            strat = strat["type"] if isinstance(strat, dict) else strat
            res = d.get("resource")
            return cls("peers", peers=list(d.get("peers", [])), strategy=SelectionStrategy(strat),
                       resource=DataSlice.from_wire(res) if res else None)
        if t == "scheduler":
            return cls("scheduler", peer=d["peer"], dataset=d["dataset"])
        raise MessageError(f"unknown reference type {t!r}")


class Fetch(Reference):
    """Any reference may be fetched (messages/src/lib.rs:275-332)."""

    @classmethod
    def uri(cls, value: str) -> "Fetch":
        return cls("uri", value=value)

    @classmethod
    def huggingface(cls, repository, revision=None, filenames=(), token=None) -> "Fetch":
        return cls("huggingface", repository=repository, revision=revision, filenames=list(filenames), token=token)

    @classmethod
    def data_peer(cls, peer: str, resource: DataSlice) -> "Fetch":
        return cls("peers", peers=[peer], strategy=SelectionStrategy.ONE, resource=resource)

    @classmethod
    def scheduler(cls, peer: str, dataset: str) -> "Fetch":
        return cls("scheduler", peer=peer, dataset=dataset)

    @classmethod
    def from_wire(cls, d: dict) -> "Fetch":
        r = Reference.from_wire(d)
        return cls(**r.__dict__)


class Send(Reference):
    """Only ``peers`` references can be sent to (messages/src/lib.rs:334-369)."""

    @classmethod
    def to_peers(cls, peers: List[str], strategy: SelectionStrategy = SelectionStrategy.ALL) -> "Send":
        return cls("peers", peers=list(peers), strategy=strategy)

    @classmethod
    def from_wire(cls, d: dict) -> "Send":
        r = Reference.from_wire(d)
        if r.type != "peers":
            raise MessageError("Send can only be created from Peers")
        return cls(**r.__dict__)


class Receive(Reference):
    """``peers`` with strategy All (messages/src/lib.rs:371-417)."""

    @classmethod
    def from_peers(cls, peers: List[str]) -> "Receive":
        return cls("peers", peers=list(peers), strategy=SelectionStrategy.ALL)

    def get_peers(self) -> List[str]:
        return self.peers

    @classmethod
    def from_wire(cls, d: dict) -> "Receive":
        r = Reference.from_wire(d)
        if r.type != "peers":
            raise MessageError("Receive can only be created from Peers")
        if r.strategy != SelectionStrategy.ALL:
            raise MessageError("Receive requires SelectionStrategy::All")
        return cls(**r.__dict__)


# -- model / executor config ----------------------------------------------------------------
MODEL_TYPES: Tuple[str, ...] = (
    "auto", "pretraining", "causal-lm", "masked-lm", "mask-generation", "seq2-seq-lm",
    "sequence-classification", "multiple-choice", "next-sentence-prediction", "token-classification",
    "question-answering", "text-encoding", "depth-estimation", "image-classification",
    "video-classification", "keypoint-detection", "keypoint-matching", "object-detection",
    "image-segmentation", "image-to-image", "semantic-segmentation", "instance-segmentation",
    "universal-segmentation", "zero-shot-image-classification", "zero-shot-object-detection",
    "audio-classification", "audio-frame-classification", "ctc", "speech-seq2-seq", "audio-x-vector",
    "text-to-spectrogram", "text-to-waveform", "audio-tokenization", "table-question-answering",
    "document-question-answering", "vison2-seq", "image-text-to-text", "time-series-prediction",
)
PREPROCESSOR_TYPES: Tuple[str, ...] = ("tokenizer", "feature", "image", "video", "auto")
LOSS_TYPES: Tuple[str, ...] = ("l1", "mse", "cross-entropy", "bce-with-logits", "kl-div")


@dataclass
class Model:
    task: str
    artifact: Fetch
    input_names: List[str] = field(default_factory=list)

    def __post_init__(self):
        if self.task not in MODEL_TYPES:
            raise MessageError(f"unknown model type {self.task!r}")

    def to_wire(self) -> dict:
        return {"task": self.task, "artifact": self.artifact.to_wire(), "input-names": list(self.input_names)}

    @classmethod
    def from_wire(cls, d: dict) -> "Model":
        return cls(d["task"], Fetch.from_wire(d["artifact"]), list(d.get("input-names", [])))


@dataclass
class Preprocessor:
    task: str
    artifact: Fetch
    input_names: List[str] = field(default_factory=list)

    def __post_init__(self):
        if self.task not in PREPROCESSOR_TYPES:
            raise MessageError(f"unknown preprocessor type {self.task!r}")

    def to_wire(self) -> dict:
        return {"task": self.task, "artifact": self.artifact.to_wire(), "input-names": list(self.input_names)}

    @classmethod
    def from_wire(cls, d: dict) -> "Preprocessor":
        return cls(d["task"], Fetch.from_wire(d["artifact"]), list(d.get("input-names", [])))


@dataclass
class Adam:
    learning_rate: float
    betas: Optional[Tuple[float, float]] = None
    epsilon: Optional[float] = None

    def to_wire(self) -> dict:
        return {"learning-rate": self.learning_rate, "betas": list(self.betas) if self.betas else None,
                "epsilon": self.epsilon}

    @classmethod
    def from_wire(cls, d: dict) -> "Adam":
        b = d.get("betas")
        return cls(float(d["learning-rate"]), tuple(b) if b else None, d.get("epsilon"))


@dataclass
class Nesterov:
    learning_rate: float
    momentum: float

    def to_wire(self) -> dict:
        return {"learning-rate": self.learning_rate, "momentum": self.momentum}

    @classmethod
    def from_wire(cls, d: dict) -> "Nesterov":
        return cls(float(d["learning-rate"]), float(d["momentum"]))


@dataclass
class LRScheduler:
    """``type``-tagged: cosine-with-warmup | linear-with-warmup | wsd."""

    type: str
    warmup_steps: int = 0
    training_steps: Optional[int] = None
    decay_steps: Optional[int] = None

    def to_wire(self) -> dict:
        d = {"type": self.type, "warmup_steps": self.warmup_steps}
        if self.type == "wsd":
            d["decay_steps"] = self.decay_steps
        else:
            d["training_steps"] = self.training_steps
        return d

    @classmethod
    def from_wire(cls, d: dict) -> "LRScheduler":
        if d["type"] not in ("cosine-with-warmup", "linear-with-warmup", "wsd"):
            raise MessageError(f"unknown scheduler type {d['type']!r}")
        return cls(d["type"], int(d.get("warmup_steps", 0)), d.get("training_steps"),
                   d.get("decay_steps", d.get("decay_step")))


@dataclass
class Loss:
    type: str

    def __post_init__(self):
        if self.type not in LOSS_TYPES:
            raise MessageError(f"unknown loss {self.type!r}")

    def to_wire(self) -> dict:
        return {"type": self.type}

    @classmethod
    def from_wire(cls, d: dict) -> "Loss":
        return cls(d["type"])


@dataclass
class TrainExecutorConfig:
    model: Model
    data: Fetch
    updates: Send
    results: Receive
    optimizer: Adam
    batch_size: int
    preprocessor: Optional[Preprocessor] = None
    scheduler: Optional[LRScheduler] = None
    # B200 extension (absent in the reference): when set, the train workers of the job synchronise
    # among themselves -- fused peer-memory outer step on GPUs, gloo all-reduce on CPU -- instead of
    # shipping files to a parameter server.  {"rank", "world", "rendezvous", "outer": Nesterov wire}
    collective: Optional[dict] = None
    # B200 extension: a replacement worker dispatched into a running parameter-server job.  It first waits
    # for the aggregate worker's cumulative update U (theta = theta_0(seed) + U), then trains normally.
    join: Optional[dict] = None
    # B200 extension: "samples" -> the executor reports the samples it processed with its pseudo-gradient and the
    # mean is taken over samples instead of over workers (absent / "uniform": the reference's 1/N intent)
    weighting: Optional[str] = None
    # B200 extension: GEMM precision of the inner steps on GPU executors ("mxfp8" | "fp8"; absent / "bf16": bf16)
    precision: Optional[str] = None
    # B200 extension: gradient accumulation inside an inner step, {"micro_batches": k, "main_grad": "fp32" | "bf16"}
    # (absent: one micro-batch per step, as the reference)
    accumulation: Optional[dict] = None

    def __post_init__(self):
        if self.accumulation is not None:
            k, mg = self.accumulation.get("micro_batches", 1), self.accumulation.get("main_grad", "fp32")
            if not isinstance(k, int) or isinstance(k, bool) or k < 1 or mg not in ("fp32", "bf16"):
                raise MessageError(f"bad accumulation spec {self.accumulation!r}")
            self.accumulation = {"micro_batches": k, "main_grad": mg} if k > 1 else None

    def to_wire(self) -> dict:
        d = {"model": self.model.to_wire(), "data": self.data.to_wire(), "updates": self.updates.to_wire(),
             "results": self.results.to_wire(), "optimizer": self.optimizer.to_wire(),
             "batch_size": self.batch_size}
        if self.preprocessor is not None:
            d["preprocessor"] = self.preprocessor.to_wire()
        if self.scheduler is not None:
            d["scheduler"] = self.scheduler.to_wire()
        if self.collective is not None:
            d["collective"] = dict(self.collective)
        if self.join is not None:
            d["join"] = dict(self.join)
        if self.weighting and self.weighting != "uniform":
            d["weighting"] = self.weighting
        if self.precision and self.precision != "bf16":
            d["precision"] = self.precision
        if self.accumulation is not None:
            d["accumulation"] = dict(self.accumulation)
        return d

    @classmethod
    def from_wire(cls, d: dict) -> "TrainExecutorConfig":
        return cls(Model.from_wire(d["model"]), Fetch.from_wire(d["data"]), Send.from_wire(d["updates"]),
                   Receive.from_wire(d["results"]), Adam.from_wire(d["optimizer"]), int(d["batch_size"]),
                   Preprocessor.from_wire(d["preprocessor"]) if d.get("preprocessor") else None,
                   LRScheduler.from_wire(d["scheduler"]) if d.get("scheduler") else None,
                   dict(d["collective"]) if d.get("collective") else None,
                   dict(d["join"]) if d.get("join") else None,
                   str(d["weighting"]) if d.get("weighting") else None,
                   str(d["precision"]) if d.get("precision") else None,
                   dict(d["accumulation"]) if d.get("accumulation") else None)


@dataclass
class AggregateExecutorConfig:
    updates: Receive
    results: Send
    optimizer: Nesterov
    # B200 extension (the reference has no checkpointing): {"dir", "every", "resume_round"?}
    checkpoint: Optional[dict] = None

    def to_wire(self) -> dict:
        d = {"updates": self.updates.to_wire(), "results": self.results.to_wire(),
             "optimizer": self.optimizer.to_wire()}
        if self.checkpoint is not None:
            d["checkpoint"] = dict(self.checkpoint)
        return d

    @classmethod
    def from_wire(cls, d: dict) -> "AggregateExecutorConfig":
        return cls(Receive.from_wire(d["updates"]), Send.from_wire(d["results"]), Nesterov.from_wire(d["optimizer"]),
                   dict(d["checkpoint"]) if d.get("checkpoint") else None)


@dataclass(frozen=True)
class ExecutorDescriptor:
    """``class``-tagged (train | aggregate) + name."""

    cls: str
    name: str

    def to_wire(self) -> dict:
        return {"class": self.cls, "name": self.name}

    @classmethod
    def from_wire(cls, d: dict) -> "ExecutorDescriptor":
        if d["class"] not in ("train", "aggregate"):
            raise MessageError(f"unknown executor class {d['class']!r}")
        return cls(d["class"], d["name"])


@dataclass
class Executor:
    """``class``-tagged: {"class": "train", "descriptor": {"name": ..}, "config": {...}}."""

    descriptor: ExecutorDescriptor
    config: Union[TrainExecutorConfig, AggregateExecutorConfig]

    @classmethod
    def train(cls, name: str, config: TrainExecutorConfig) -> "Executor":
        return cls(ExecutorDescriptor("train", name), config)

    @classmethod
    def aggregate(cls, name: str, config: AggregateExecutorConfig) -> "Executor":
        return cls(ExecutorDescriptor("aggregate", name), config)

    def to_wire(self) -> dict:
        return {"class": self.descriptor.cls, "descriptor": {"name": self.descriptor.name},
                "config": self.config.to_wire()}

    @classmethod
    def from_wire(cls, d: dict) -> "Executor":
        k = d["class"]
        desc = ExecutorDescriptor(k, d["descriptor"]["name"])
        if k == "train":
            return cls(desc, TrainExecutorConfig.from_wire(d["config"]))
        if k == "aggregate":
            return cls(desc, AggregateExecutorConfig.from_wire(d["config"]))
        raise MessageError(f"unknown executor class {k!r}")


@dataclass
class JobSpec:
    job_id: str
    executor: Executor

    def to_wire(self) -> dict:
        return {"job_id": self.job_id, "executor": self.executor.to_wire()}

    @classmethod
    def from_wire(cls, d: dict) -> "JobSpec":
        return cls(str(d["job_id"]), Executor.from_wire(d["executor"]))

    def to_json(self) -> str:
        return json.dumps(self.to_wire())


@dataclass
class WorkerSpec:
    resources: Resources
    executor: List[ExecutorDescriptor] = field(default_factory=list)

    def to_wire(self) -> dict:
        return {"resources": self.resources.to_wire(), "executor": [e.to_wire() for e in self.executor]}

    @classmethod
    def from_wire(cls, d: dict) -> "WorkerSpec":
        return cls(Resources.from_wire(d["resources"]), [ExecutorDescriptor.from_wire(e) for e in d.get("executor", [])])


class JobStatus(str, Enum):
    RUNNING = "Running"
    FINISHED = "Finished"
    FAILED = "Failed"
    UNKNOWN = "Unknown"


@dataclass
class ArtifactHeader:
    job_id: str
    epoch: int

    def to_wire(self) -> dict:
        return {"job_id": self.job_id, "epoch": self.epoch}


@dataclass
class ParameterStreamHeader:
    stream_id: str
    data_size: int

    def to_wire(self) -> dict:
        return {"stream_id": self.stream_id, "data_size": self.data_size}


# -- request_worker (gossip payload) -----------------------------------------------------------
@dataclass
class RequestWorker:
    id: str
    spec: WorkerSpec
    timeout: float
    bid: float

    def to_wire(self) -> dict:
        return {"id": self.id, "spec": self.spec.to_wire(), "timeout": time_to_wire(self.timeout), "bid": self.bid}

    @classmethod
    def from_wire(cls, d: dict) -> "RequestWorker":
        return cls(str(d["id"]), WorkerSpec.from_wire(d["spec"]), time_from_wire(d["timeout"]), float(d["bid"]))


# -- /hypha-api/0.0.1 ------------------------------------------------------------------------------
@dataclass
class ApiRequest:
    """Externally tagged: {"WorkerOffer": {...}} etc. ``kind`` is the variant name."""

    kind: str
    body: Dict[str, Any]

    # "JobMembership" is a B200 extension: the scheduler tells an aggregate job which train workers are
    # still part of the job after a lease was lost (the reference has no such message: its PS waits forever)
    KINDS = ("WorkerOffer", "RenewLease", "JobStatus", "DispatchJob", "ParameterPull", "ParameterPush", "Data",
             "JobMembership")

    def to_wire(self) -> dict:
        return {self.kind: self.body}

    @classmethod
    def from_wire(cls, d: dict) -> "ApiRequest":
        (k, v), = d.items()
        if k not in cls.KINDS:
            raise MessageError(f"unknown api request {k!r}")
        return cls(k, v)

    # constructors ------------------------------------------------------------
    @classmethod
    def worker_offer(cls, id: str, request_id: str, price: float, resources: Resources, timeout: float):
        return cls("WorkerOffer", {"id": id, "request_id": request_id, "price": price,
                                   "resources": resources.to_wire(), "timeout": time_to_wire(timeout)})

    @classmethod
    def renew_lease(cls, id: str):
        return cls("RenewLease", {"id": id})

    @classmethod
    def dispatch_job(cls, id: str, spec: JobSpec):
        return cls("DispatchJob", {"id": id, "spec": spec.to_wire()})

    @classmethod
    def job_status(cls, task_id: str, status: JobStatus):
        return cls("JobStatus", {"task_id": task_id, "status": {"type": status.value}})

    @classmethod
    def data(cls, dataset: str):
        return cls("Data", {"dataset": dataset})

    @classmethod
    def job_membership(cls, job_id: str, workers: List[str]):
        return cls("JobMembership", {"job_id": job_id, "workers": list(workers)})

    @classmethod
    def parameter_pull(cls, job_id: str, key: str, version: Optional[int] = None):
        return cls("ParameterPull", {"job_id": job_id, "key": key, "version": version})

    @classmethod
    def parameter_push(cls, job_id: str, key: str, data_stream_id: str, data_size: int, version: Optional[int] = None):
        return cls("ParameterPush", {"job_id": job_id, "key": key, "version": version,
                                     "data_stream_id": data_stream_id, "data_size": data_size})


@dataclass
class ApiResponse:
    kind: str
    body: Any

    def to_wire(self) -> dict:
        return {self.kind: self.body}

    @classmethod
    def from_wire(cls, d: dict) -> "ApiResponse":
        (k, v), = d.items()
        return cls(k, v)

    @classmethod
    def worker_offer(cls):
        return cls("WorkerOffer", {})

    @classmethod
    def renewed(cls, id: str, timeout: float):
        return cls("RenewLease", {"Renewed": {"id": id, "timeout": time_to_wire(timeout)}})

    @classmethod
    def renew_failed(cls):
        return cls("RenewLease", "Failed")

    @classmethod
    def dispatched(cls, id: str, timeout: float):
        return cls("DispatchJob", {"Dispatched": {"id": id, "timeout": time_to_wire(timeout)}})

    @classmethod
    def dispatch_failed(cls):
        return cls("DispatchJob", "Failed")

    @classmethod
    def job_status(cls):
        return cls("JobStatus", {})

    @classmethod
    def job_membership(cls, known: bool):
        return cls("JobMembership", {"known": bool(known)})

    @classmethod
    def data_success(cls, data_provider: str, index: int):
        return cls("Data", {"Success": {"data_provider": data_provider, "index": index}})

    @classmethod
    def data_not_found(cls):
        return cls("Data", "NotFound")

    @classmethod
    def data_error(cls, msg: str):
        return cls("Data", {"Error": msg})

    def variant(self) -> Tuple[str, Any]:
        """('Renewed', {...}) / ('Failed', None) for nested enum bodies."""
        if isinstance(self.body, str):
            return self.body, None
        if isinstance(self.body, dict) and len(self.body) == 1:
            (k, v), = self.body.items()
            return k, v
        return "", self.body


# -- /hypha-health/0.0.1 --------------------------------------------------------------------------------
@dataclass
class HealthRequest:
    def to_wire(self) -> dict:
        return {}


@dataclass
class HealthResponse:
    healthy: bool

    def to_wire(self) -> dict:
        return {"healthy": self.healthy}

    @classmethod
    def from_wire(cls, d: dict) -> "HealthResponse":
        return cls(bool(d["healthy"]))


# -- /hypha-progress/0.0.1 ---------------------------------------------------------------------------------
@dataclass
class Progress:
    """kebab-case externally tagged: "update" | "updated" | "update-received" |
    {"status": {"batch_size": n}} | {"metrics": {"round": r, "metrics": {...}}}."""

    kind: str
    batch_size: int = 0
    round: int = 0
    metrics: Dict[str, float] = field(default_factory=dict)

    KINDS = ("status", "metrics", "update", "updated", "update-received")

    def to_wire(self):
        if self.kind == "status":
            return {"status": {"batch_size": self.batch_size}}
        if self.kind == "metrics":
            return {"metrics": {"round": self.round, "metrics": dict(self.metrics)}}
        return self.kind

    @classmethod
    def from_wire(cls, d) -> "Progress":
        if isinstance(d, str):
            if d not in ("update", "updated", "update-received"):
                raise MessageError(f"unknown progress {d!r}")
            return cls(d)
        (k, v), = d.items()
        if k == "status":
            return cls("status", batch_size=int(v["batch_size"]))   # extra keys (round_time) ignored, quirk Q2
        if k == "metrics":
            return cls("metrics", round=int(v["round"]), metrics={a: float(b) for a, b in v["metrics"].items()})
        if k in ("update", "updated", "update-received"):
            return cls(k)
        raise MessageError(f"unknown progress {k!r}")


@dataclass
class ProgressRequest:
    job_id: str
    progress: Progress

    def to_wire(self) -> dict:
        return {"job_id": self.job_id, "progress": self.progress.to_wire()}

    @classmethod
    def from_wire(cls, d: dict) -> "ProgressRequest":
        return cls(str(d["job_id"]), Progress.from_wire(d["progress"]))


@dataclass
class ProgressResponse:
    """``type``-tagged: Ok | Continue | ScheduleUpdate{counter} | Done | Error."""

    type: str
    counter: int = 0
    # B200 extension, only on the answer to `update` of a collective job: who takes part in this outer step
    # {"generation": g, "members": [peer ids in rank order], "rendezvous": "tcp://..."}
    membership: Optional[dict] = None

    def to_wire(self) -> dict:
        d = {"type": self.type}
        if self.type == "ScheduleUpdate":
            d["counter"] = self.counter
        if self.membership is not None:
            d["membership"] = dict(self.membership)
        return d

    @classmethod
    def from_wire(cls, d: dict) -> "ProgressResponse":
        if d["type"] not in ("Ok", "Continue", "ScheduleUpdate", "Done", "Error"):
            raise MessageError(f"unknown progress response {d['type']!r}")
        return cls(d["type"], int(d.get("counter", 0)), dict(d["membership"]) if d.get("membership") else None)


# -- codecs -----------------------------------------------------------------------------------------------
def encode_cbor(obj: Any) -> bytes:
    w = obj.to_wire() if hasattr(obj, "to_wire") else obj
    return cbor2.dumps(w)


def decode_cbor(data: bytes) -> Any:
    return cbor2.loads(data)


def encode_json(obj: Any) -> str:
    w = obj.to_wire() if hasattr(obj, "to_wire") else obj
    return json.dumps(w)


def encode_pull_header(resource: DataSlice) -> bytes:
    """Pull-stream header: u64 LE length (< 1 MiB) + JSON (crates/network/src/stream_pull.rs:28,128-146)."""
    body = json.dumps(resource.to_wire()).encode()
    if len(body) >= 1 << 20:
        raise MessageError("pull header too large")
    return len(body).to_bytes(8, "little") + body


def decode_pull_header(stream: io.BufferedIOBase) -> DataSlice:
    raw = stream.read(8)
    if len(raw) != 8:
        raise MessageError("truncated pull header")
    n = int.from_bytes(raw, "little")
    if n >= 1 << 20:
        raise MessageError("pull header too large")
    body = stream.read(n)
    if len(body) != n:
        raise MessageError("truncated pull header body")
    return DataSlice.from_wire(json.loads(body))
