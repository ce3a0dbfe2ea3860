"""Scheduler role: allocation, dispatch, progress/data scheduling, metrics bridge."""
from .allocator import (AllocatorError, GreedyWorkerAllocator, LeaseExpired, NoOffersReceived, Offer, Task,
                        TaskError, WorkerHandle)
from .scheduling import (AimConnector, BatchSchedulerService, Connector, DataSchedulerService, FanOutConnector,
                         JsonlConnector, MemoryConnector, MetricsBridge, NoOpConnector, OtelConnector)

from .job import DiLoCoJobRun, JobError, JobOutcome, worker_batch_size

__all__ = ["DiLoCoJobRun", "JobError", "JobOutcome", "worker_batch_size", "AllocatorError", "GreedyWorkerAllocator", "LeaseExpired", "NoOffersReceived", "Offer", "Task",
           "TaskError", "WorkerHandle", "AimConnector", "BatchSchedulerService", "Connector",
           "DataSchedulerService", "FanOutConnector", "JsonlConnector", "MemoryConnector", "MetricsBridge", "NoOpConnector",
           "OtelConnector"]
