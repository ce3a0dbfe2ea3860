"""Worker role: arbiter (offers/leases/dispatch), job manager + executors, bridge, connector."""
from .arbiter import Arbiter
from .bridge import Bridge, BridgeError, Connector, InvalidPath, InvalidResource, safe_join
from .jobs import (Execution, Job, JobManager, JobManagerError, ParameterServerExecutor, ProcessExecutor,
                   nesterov_step, substitute_placeholders)

__all__ = ["Arbiter", "Bridge", "BridgeError", "Connector", "InvalidPath", "InvalidResource", "safe_join",
           "Execution", "Job", "JobManager", "JobManagerError", "ParameterServerExecutor", "ProcessExecutor",
           "nesterov_step", "substitute_placeholders"]
