"""Resource monitor of a worker (the reference only drafts one: rfc/decentralized_resource_allocation_protocol.md:
1392-1423): NVML samples of the GPUs a worker offers -- utilisation, memory, temperature, power, SM clock and the
throttle reasons -- exported as OpenTelemetry gauges (``hypha.gpu.*`` with a ``gpu`` attribute) next to the
bandwidth counters, and a health summary the worker logs when a device is being slowed down by the hardware.

A B200 under the training step sits at its power cap (``sw_power_cap``: normal, `bench.py` reports it); what an
operator must hear about is ``hw_slowdown`` / ``hw_thermal_slowdown`` / ``sw_thermal_slowdown`` -- a step that is
20 % slower on one GPU stalls every round of a DiLoCo job by that much.
"""
from __future__ import annotations

import threading
import time
from typing import Callable, Dict, List, Optional, Sequence

# NVML clocks-event (throttle) reason bits
REASONS = {
    0x0000000000000001: "gpu_idle", 0x0000000000000002: "applications_clocks_setting", 0x0000000000000004: "sw_power_cap",
    0x0000000000000008: "hw_slowdown", 0x0000000000000010: "sync_boost", 0x0000000000000020: "sw_thermal_slowdown",
    0x0000000000000040: "hw_thermal_slowdown", 0x0000000000000080: "hw_power_brake_slowdown",
    0x0000000000000100: "display_clock_setting",
}
ALARMING = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "hw_power_brake_slowdown")


def decode_reasons(mask: int) -> List[str]:
    return [name for bit, name in REASONS.items() if mask & bit]


class GpuMonitor:
    """``nvml``: a module with the pynvml API (injected in tests); ``devices``: NVML indices to watch."""

    def __init__(self, devices: Optional[Sequence[int]] = None, nvml=None, interval_s: float = 5.0,
                 log: Optional[Callable[[str], None]] = None):
        if nvml is None:
            import pynvml as nvml                       # nvidia-ml-py
        self.nvml = nvml
        self.interval_s = interval_s
        self.log = log or (lambda m: None)
        nvml.nvmlInit()
        count = int(nvml.nvmlDeviceGetCount())
        self.devices = [int(d) for d in (devices if devices is not None else range(count)) if 0 <= int(d) < count]
        self._handles = {d: nvml.nvmlDeviceGetHandleByIndex(d) for d in self.devices}
        self._last: Dict[int, dict] = {}
        self._lock = threading.Lock()
        self._stop = threading.Event()
        self._thread: Optional[threading.Thread] = None
        self._alarmed: Dict[int, tuple] = {}

    # -- sampling ---------------------------------------------------------------------------------------------------
    def _try(self, fn, *args, default=None):
        try:
            return fn(*args)
        except Exception:               # noqa: BLE001 - a query a driver / board does not support is just absent
            return default

    def sample(self) -> Dict[int, dict]:
        n = self.nvml
        out: Dict[int, dict] = {}
        for d, h in self._handles.items():
            util = self._try(n.nvmlDeviceGetUtilizationRates, h)
            mem = self._try(n.nvmlDeviceGetMemoryInfo, h)
            temp = self._try(n.nvmlDeviceGetTemperature, h, getattr(n, "NVML_TEMPERATURE_GPU", 0))
            power = self._try(n.nvmlDeviceGetPowerUsage, h)                      # milliwatts
            clock = self._try(n.nvmlDeviceGetClockInfo, h, getattr(n, "NVML_CLOCK_SM", 1))
            getter = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                getattr(n, "nvmlDeviceGetCurrentClocksThrottleReasons", None)
            mask = self._try(getter, h, default=0) if getter is not None else 0
            out[d] = {"utilization_pct": float(getattr(util, "gpu", 0) or 0) if util is not None else None,
                      "memory_used": int(mem.used) if mem is not None else None,
                      "memory_total": int(mem.total) if mem is not None else None,
                      "temperature_c": float(temp) if temp is not None else None,
                      "power_w": float(power) / 1e3 if power is not None else None,
                      "sm_clock_mhz": float(clock) if clock is not None else None,
                      "throttle_reasons": decode_reasons(int(mask or 0)), "t": time.time()}
        with self._lock:
            self._last = out
        return out

    def last(self) -> Dict[int, dict]:
        with self._lock:
            return dict(self._last)

    def health(self) -> dict:
        """``{"healthy": bool, "throttled": {gpu: [alarming reasons]}}`` from the newest sample."""
        bad = {d: [r for r in s["throttle_reasons"] if r in ALARMING] for d, s in self.last().items()}
        bad = {d: r for d, r in bad.items() if r}
        return {"healthy": not bad, "throttled": bad}

    # -- export ------------------------------------------------------------------------------------------------------
    def register(self, telemetry) -> None:
        """Observable gauges on the role's meter; every collection reads the newest sample (no NVML call on the
        exporter's thread)."""
        from opentelemetry.metrics import Observation

        def gauge(name: str, key: str, unit: str):
            def cb(_options):
                return [Observation(s[key], {"gpu": str(d)}) for d, s in self.last().items() if s.get(key) is not None]

            telemetry.meter.create_observable_gauge(name, callbacks=[cb], unit=unit)

        gauge("hypha.gpu.utilization", "utilization_pct", "%")
        gauge("hypha.gpu.memory.used", "memory_used", "By")
        gauge("hypha.gpu.temperature", "temperature_c", "Cel")
        gauge("hypha.gpu.power", "power_w", "W")
        gauge("hypha.gpu.clock.sm", "sm_clock_mhz", "MHz")

        def throttle(_options):
            return [Observation(int(bool([r for r in s["throttle_reasons"] if r in ALARMING])), {"gpu": str(d)})
                    for d, s in self.last().items()]

        telemetry.meter.create_observable_gauge("hypha.gpu.throttled", callbacks=[throttle], unit="1")

    # -- background loop ------------------------------------------------------------------------------------------------
    def _tick(self) -> None:
        self.sample()
        for d, reasons in self.health()["throttled"].items():
            key = tuple(reasons)
            if self._alarmed.get(d) != key:              # say it once per change, not every 5 s
                self._alarmed[d] = key
                self.log(f"gpu {d} is being slowed down by the hardware: {', '.join(reasons)} "
                         f"({self.last()[d].get('temperature_c')} C, {self.last()[d].get('sm_clock_mhz')} MHz)")
        for d in list(self._alarmed):
            if d not in self.health()["throttled"]:
                self._alarmed.pop(d)
                self.log(f"gpu {d} runs at its normal clocks again")

    def start(self) -> None:
        if self._thread is not None or not self.devices:
            return

        def loop():
            while not self._stop.is_set():
                try:
                    self._tick()
                except Exception:       # noqa: BLE001 - monitoring never takes the worker down
                    pass
                self._stop.wait(self.interval_s)

        self._thread = threading.Thread(target=loop, name="hypha-gpu-monitor", daemon=True)
        self._thread.start()

    def stop(self) -> None:
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2.0)
            self._thread = None
        try:
            self.nvml.nvmlShutdown()
        except Exception:               # noqa: BLE001
            pass


def start_for_worker(cfg, telemetry, log: Callable[[str], None]) -> Optional[GpuMonitor]:
    """The monitor of a worker role: the device it is pinned to (`device`), else every visible GPU.  None on a
    machine without NVML / GPUs or when ``HYPHA_GPU_MONITOR=0``."""
    import os

    if os.environ.get("HYPHA_GPU_MONITOR", "1") == "0" or float(getattr(cfg.resources, "gpu", 0) or 0) <= 0:
        return None
    try:
        dev = getattr(cfg, "device", None)
        mon = GpuMonitor([dev] if dev is not None else None,
                         interval_s=float(os.environ.get("HYPHA_GPU_MONITOR_S", "5")), log=log)
    except Exception:                   # noqa: BLE001 - no driver, no library: nothing to monitor
        return None
    if not mon.devices:
        return None
    mon.register(telemetry)
    mon.start()
    return mon
