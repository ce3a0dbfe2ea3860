"""Scheduler role: allocation, dispatch, progress/data scheduling, metrics bridge."""
from .allocator import (AllocatorError, GreedyWorkerAllocator, LeaseExpired, NoOffersReceived, Offer, Task,
                        TaskError, WorkerHandle)
from .scheduling import (AimConnector, BatchSchedulerService, Connector, DataSchedulerService, JsonlConnector,
                         MemoryConnector, MetricsBridge, NoOpConnector)

__all__ = ["AllocatorError", "GreedyWorkerAllocator", "LeaseExpired", "NoOffersReceived", "Offer", "Task",
           "TaskError", "WorkerHandle", "AimConnector", "BatchSchedulerService", "Connector",
           "DataSchedulerService", "JsonlConnector", "MemoryConnector", "MetricsBridge", "NoOpConnector"]
