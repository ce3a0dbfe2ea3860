"""AIM metrics driver: ``POST /status`` -> ``aim.Run.track`` (reference: drivers/aim-driver/main.py:1-13).

``aim`` is optional; without it metrics land in a JSONL file so the scheduler's ``status_bridge``
target still works offline.  Run: ``uvicorn hypha_b200.drivers.aim_driver:app --port 61000``."""
from __future__ import annotations

import json
import os
import threading

from fastapi import FastAPI
from pydantic import BaseModel

app = FastAPI(title="hypha aim driver")
_lock = threading.Lock()
_run = None
JSONL = os.environ.get("HYPHA_AIM_JSONL", "hypha-metrics.jsonl")


class AimMetrics(BaseModel):
    worker_id: str
    round: int
    metric_name: str
    value: float


def _track(m: AimMetrics) -> None:
    global _run
    try:
        import aim  # type: ignore

        if _run is None:
            _run = aim.Run()
        _run.track(m.value, name=f"{m.worker_id}_{m.metric_name}", epoch=m.round)
    except ImportError:
        with _lock, open(JSONL, "a") as f:
            f.write(json.dumps(m.model_dump()) + "\n")


@app.post("/status")
def status(m: AimMetrics) -> dict:
    _track(m)
    return {}
