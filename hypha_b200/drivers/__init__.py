"""Optional sidecar services next to the roles.

`aim_driver` -- the metrics sink the scheduler's `AimConnector` POSTs to (reference: drivers/aim-driver/main.py:1-13):
``POST /status`` with ``{"worker_id", "round", "metric_name", "value"}`` -> ``aim.Run.track(value, name=f"{worker_id}_
{metric_name}", epoch=round)``; without the `aim` package the records go to a JSONL file (``HYPHA_AIM_JSONL``), so a
scheduler's ``status_bridge`` target works on a box without AIM.  Run it with
``uvicorn hypha_b200.drivers.aim_driver:app --port 61000``; the module is imported on demand (it needs `fastapi`).
"""
