"""Job execution on a worker: JobManager, the process executor, the parameter-server executor.

Reference: crates/worker/src/job_manager.rs:36-211, crates/worker/src/executor/process.rs:77-205
(work dir ``hypha-{uuid}``, ``bridge.sock``, placeholder substitution, SIGTERM -> 5 s -> SIGKILL,
cleanup), crates/worker/src/executor/parameter_server.rs:73-446 (receive N updates, average,
Nesterov, broadcast, ``Updated``).

Differences by design: the parameter server takes a TRUE mean over the N updates (the reference's
pairwise running average is only right for N=2), has a receive timeout with
membership shrink instead of waiting forever (quirk Q8), and jobs are keyed by lease id (quirk
Q7).  On an NVSwitch box the aggregate role is normally subsumed by the fused peer-memory outer
step (hypha_b200/parallel/diloco.py); this executor keeps the file-based path for the CPU
plumbing configuration and for API parity.
"""
from __future__ import annotations

import asyncio
import hashlib
import json
import os
import shutil
import time
import signal
import sys
import uuid
from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional

from .. import messages as M
from ..transport import Network, NetworkError
from .bridge import Bridge, Connector
from .ps_round import SafetensorsFormatError, outer_round, read_header

KILL_GRACE_S = 5.0


def substitute_placeholders(args: List[str], socket_path: str, work_dir: str, job_json: str) -> List[str]:
    """``{SOCKET_PATH}`` / ``{WORK_DIR}`` / ``{JOB_JSON}`` (process.rs:201-205)."""
    out = []
    for a in args:
        out.append(a.replace("{SOCKET_PATH}", socket_path).replace("{WORK_DIR}", work_dir)
                   .replace("{JOB_JSON}", job_json))
    return out


class Execution:
    """A running job: ``wait()`` for the exit status, ``cancel()`` to stop it."""

    def __init__(self) -> None:
        self.done = asyncio.Event()
        self.status = M.JobStatus.RUNNING
        self._cancel = asyncio.Event()
        self.members: Optional[List[str]] = None          # set by the scheduler after a lease was lost
        self.members_changed = asyncio.Event()

    def cancel(self) -> None:
        self._cancel.set()

    def set_members(self, workers: List[str]) -> None:
        self.members = list(workers)
        self.members_changed.set()

    async def wait(self) -> M.JobStatus:
        await self.done.wait()
        return self.status


class ProcessExecutor:
    """Runs the job as a subprocess supervised over a per-job bridge socket."""

    def __init__(self, connector: Connector, network: Network, work_dir_base: str, cmd: str, args: List[str],
                 env: Optional[Dict[str, str]] = None, log: Callable[[str], None] = print):
        self.connector, self.network = connector, network
        self.base, self.cmd, self.args, self.env, self.log = work_dir_base, cmd, args, env or {}, log

    async def execute(self, spec: M.JobSpec, task_id: str, scheduler: str) -> Execution:
        ex = Execution()
        work_dir = os.path.join(self.base, f"hypha-{uuid.uuid4()}")
        os.makedirs(work_dir, exist_ok=True)
        sock = os.path.join(work_dir, "bridge.sock")
        sock_dir = None
        if len(sock.encode()) > 100:            # sun_path is 108 bytes; fall back to a short private dir
            import tempfile

            sock_dir = tempfile.mkdtemp(prefix="hb-", dir="/tmp")
            os.chmod(sock_dir, 0o700)
            sock = os.path.join(sock_dir, "bridge.sock")
        bridge = Bridge(self.connector, self.network, work_dir, sock, spec.job_id, scheduler, log=self.log)
        await bridge.start()
        job_json = spec.to_json()
        argv = substitute_placeholders(self.args, sock, work_dir, job_json)
        cmd = sys.executable if self.cmd in ("python", "python3") else self.cmd
        env = dict(os.environ)
        # executors of several workers usually share a machine: OpenMP threads that spin while waiting starve
        # the neighbours (15x the CPU time for the same CPU job when oversubscribed); override per executor
        # with `[[executors]].env`
        for k, v in (("OMP_WAIT_POLICY", "PASSIVE"), ("GOMP_SPINCOUNT", "0"), ("KMP_BLOCKTIME", "0")):
            env.setdefault(k, v)
        env.update(self.env)
        env.update({"SOCKET_PATH": sock, "WORK_DIR": work_dir, "JOB_JSON": job_json})
        try:
            proc = await asyncio.create_subprocess_exec(cmd, *argv, env=env, stdout=asyncio.subprocess.PIPE,
                                                        stderr=asyncio.subprocess.STDOUT, start_new_session=True)
        except OSError:
            await bridge.stop()
            shutil.rmtree(work_dir, ignore_errors=True)
            if sock_dir:
                shutil.rmtree(sock_dir, ignore_errors=True)
            raise
        ex.pid = proc.pid  # type: ignore[attr-defined]
        ex.work_dir = work_dir  # type: ignore[attr-defined]

        async def forward():
            assert proc.stdout is not None
            async for line in proc.stdout:
                self.log(f"[job {spec.job_id[:8]}] {line.decode(errors='replace').rstrip()}")

        async def supervise():
            fwd = asyncio.ensure_future(forward())
            waiter = asyncio.ensure_future(proc.wait())
            canceller = asyncio.ensure_future(ex._cancel.wait())
            try:
                await asyncio.wait({waiter, canceller}, return_when=asyncio.FIRST_COMPLETED)
                if not waiter.done():
                    try:
                        os.killpg(proc.pid, signal.SIGTERM)
                    except ProcessLookupError:
                        pass
                    try:
                        await asyncio.wait_for(asyncio.shield(waiter), KILL_GRACE_S)
                    except asyncio.TimeoutError:
                        try:
                            os.killpg(proc.pid, signal.SIGKILL)
                        except ProcessLookupError:
                            pass
                        await waiter
                    ex.status = M.JobStatus.FAILED
                else:
                    ex.status = M.JobStatus.FINISHED if proc.returncode == 0 else M.JobStatus.FAILED
            finally:
                canceller.cancel()
                try:
                    await asyncio.wait_for(fwd, 1.0)
                except (asyncio.TimeoutError, Exception):
                    fwd.cancel()
                await bridge.stop()
                shutil.rmtree(work_dir, ignore_errors=True)
                if sock_dir:
                    shutil.rmtree(sock_dir, ignore_errors=True)
                ex.done.set()

        ex._task = asyncio.ensure_future(supervise())  # type: ignore[attr-defined]
        return ex


# ---------------------------------------------------------------------------------------------
# parameter server
# ---------------------------------------------------------------------------------------------
def nesterov_step(avg: Dict[str, Any], momentum: Optional[Dict[str, Any]], lr: float, mu: float):
    """``m = g`` (first round) else ``mu*m + g``; ``u = lr*(mu*m + g)``.
    Vector (parameter_server.rs:452-524): lr=.1 mu=.7 g0=.5 -> u0=.085;
    g1=[.1,.2,.3,.4,.5] -> u1=[.0415,.0585,.0755,.0925,.1095]."""
    new_m, upd = {}, {}
    for k, g in avg.items():
        m = g.clone() if momentum is None else momentum[k] * mu + g
        new_m[k] = m
        upd[k] = (m * mu + g) * lr
    return new_m, upd


def latest_checkpoint(ckpt_dir: str) -> Optional[dict]:
    """``{"round": r, "path": <dir>/round-r}`` of the newest complete job checkpoint, or None."""
    try:
        meta = json.load(open(os.path.join(ckpt_dir, "latest.json")))
        path = os.path.join(ckpt_dir, f"round-{int(meta['round'])}")
        if os.path.isfile(os.path.join(path, "cumulative_update")) and os.path.isfile(os.path.join(path, "momentum")):
            return {"round": int(meta["round"]), "path": path, "kind": "parameter-server"}
        if os.path.isfile(os.path.join(path, "collective_state")):
            return {"round": int(meta["round"]), "path": path, "kind": "collective"}
        if os.path.isfile(os.path.join(path, "collective_params")):     # fused GPU engine: params + one outer shard per rank
            return {"round": int(meta["round"]), "path": path, "kind": "collective"}
    except (OSError, ValueError, KeyError):
        pass
    return None


def save_job_checkpoint(ckpt_dir: str, round_no: int, cum_path: str, momentum_path: str, keep: int = 2) -> str:
    """Copy the aggregate job's outer state (cumulative update U + momentum) to ``<dir>/round-<n>`` and flip
    ``latest.json`` atomically; older checkpoints beyond ``keep`` are removed."""
    return publish_checkpoint(ckpt_dir, round_no, {"cumulative_update": cum_path, "momentum": momentum_path}, keep)


def publish_sharded_checkpoint(ckpt_dir: str, round_no: int, name: str, src: str, *, expect: Optional[List[str]] = None,
                               wait_s: float = 120.0, keep: int = 2) -> Optional[str]:
    """Several processes of one box write ONE checkpoint: each moves its own file into ``round-<n>/`` atomically
    (``name.tmp`` -> ``name``).  The process that passes ``expect`` (the complete list of file names: rank 0)
    waits until all of them exist, flips ``latest.json`` -- the round is never announced half-written -- and prunes
    old rounds.  Returns the round directory for that process, None for the others."""
    dst = os.path.join(ckpt_dir, f"round-{round_no}")
    os.makedirs(dst, exist_ok=True)
    tmp = os.path.join(dst, name + ".tmp")
    shutil.copyfile(src, tmp)
    os.replace(tmp, os.path.join(dst, name))
    if expect is None:
        return None
    deadline = time.monotonic() + wait_s
    while not all(os.path.isfile(os.path.join(dst, n)) for n in expect):
        if time.monotonic() > deadline:
            missing = [n for n in expect if not os.path.isfile(os.path.join(dst, n))]
            raise TimeoutError(f"checkpoint of round {round_no}: still missing {missing}")
        time.sleep(0.05)
    meta_tmp = os.path.join(ckpt_dir, "latest.json.tmp")
    json.dump({"round": round_no}, open(meta_tmp, "w"))
    os.replace(meta_tmp, os.path.join(ckpt_dir, "latest.json"))
    rounds = sorted(int(d.split("-")[1]) for d in os.listdir(ckpt_dir) if d.startswith("round-") and d.split("-")[1].isdigit())
    for r in rounds[:-keep]:
        shutil.rmtree(os.path.join(ckpt_dir, f"round-{r}"), ignore_errors=True)
    return dst


def publish_checkpoint(ckpt_dir: str, round_no: int, files: Dict[str, str], keep: int = 2) -> str:
    """``files`` = {name in the checkpoint: path of the file to copy}.  The round directory appears complete or
    not at all, ``latest.json`` flips last."""
    dst = os.path.join(ckpt_dir, f"round-{round_no}")
    tmp = dst + ".tmp"
    shutil.rmtree(tmp, ignore_errors=True)
    os.makedirs(tmp, exist_ok=True)
    for name, src in files.items():
        shutil.copyfile(src, os.path.join(tmp, name))
    shutil.rmtree(dst, ignore_errors=True)
    os.replace(tmp, dst)
    meta_tmp = os.path.join(ckpt_dir, "latest.json.tmp")
    json.dump({"round": round_no}, open(meta_tmp, "w"))
    os.replace(meta_tmp, os.path.join(ckpt_dir, "latest.json"))
    rounds = sorted(int(d.split("-")[1]) for d in os.listdir(ckpt_dir) if d.startswith("round-") and d.split("-")[1].isdigit())
    for r in rounds[:-keep]:
        shutil.rmtree(os.path.join(ckpt_dir, f"round-{r}"), ignore_errors=True)
    return dst


class ParameterServerExecutor:
    def __init__(self, connector: Connector, network: Network, work_dir_base: str,
                 receive_timeout_s: float = 600.0, log: Callable[[str], None] = print):
        self.connector, self.network, self.base = connector, network, work_dir_base
        self.receive_timeout_s, self.log = receive_timeout_s, log

    async def execute(self, spec: M.JobSpec, task_id: str, scheduler: str) -> Execution:
        cfg = spec.executor.config
        assert isinstance(cfg, M.AggregateExecutorConfig)
        ex = Execution()
        work_dir = os.path.join(self.base, f"hypha-{uuid.uuid4()}")
        os.makedirs(work_dir, exist_ok=True)
        ex.work_dir = work_dir  # type: ignore[attr-defined]

        async def run():
            from safetensors.torch import save_file

            workers = list(cfg.updates.get_peers())
            allowed = set(workers)                 # live allow-list of the push-stream receiver
            result_peers = list(cfg.results.peers)
            pending_join: List[str] = []
            cum_path = os.path.join(work_dir, "cumulative_update")     # U = sum of all outer updates so far
            rnd = 0
            ckpt = cfg.checkpoint or {}
            ckpt_dir, ckpt_every = ckpt.get("dir"), max(1, int(ckpt.get("every", 1)))
            if ckpt_dir and ckpt.get("resume_round"):
                # resume: outer state from the checkpoint; every worker is (re)admitted like a newcomer and
                # receives U at this first round boundary
                found = latest_checkpoint(ckpt_dir)
                if found is None or found["round"] != int(ckpt["resume_round"]):
                    raise RuntimeError(f"checkpoint of round {ckpt['resume_round']} not found in {ckpt_dir}")
                shutil.copyfile(os.path.join(found["path"], "cumulative_update"), cum_path)
                shutil.copyfile(os.path.join(found["path"], "momentum"), os.path.join(work_dir, "momentum"))
                rnd = found["round"]
                pending_join, workers, result_peers = list(workers), [], []
                self.log(f"[ps {spec.job_id[:8]}] resumed from round {rnd} ({ckpt_dir})")
            incoming = self.connector.receive(cfg.updates, job_id=spec.job_id, allowed=allowed)
            try:
                nxt: Optional[asyncio.Future] = None
                while True:
                    # round boundary: admit replacement workers.  They start from theta_0(seed) + U, which
                    # is exactly the theta_prev every other worker holds now.
                    for w in pending_join:
                        if os.path.exists(cum_path):
                            await self.connector.send(M.Send.to_peers([w]), cum_path, job_id=spec.job_id)
                        else:
                            first = os.path.join(work_dir, "empty_update")
                            save_file({}, first)
                            await self.connector.send(M.Send.to_peers([w]), first, job_id=spec.job_id)
                            os.unlink(first)
                        workers.append(w)
                        result_peers.append(w)
                        self.log(f"[ps {spec.job_id[:8]}] round {rnd}: {w} joined (cumulative update sent)")
                    pending_join.clear()
                    got: Dict[str, str] = {}
                    while len(got) < len(workers):
                        if nxt is None:
                            nxt = asyncio.ensure_future(incoming.__anext__())
                        mem = asyncio.ensure_future(ex.members_changed.wait())
                        done, _ = await asyncio.wait({nxt, mem}, timeout=self.receive_timeout_s,
                                                     return_when=asyncio.FIRST_COMPLETED)
                        if mem not in done:
                            mem.cancel()
                        if mem in done and ex.members is not None:
                            # the scheduler lost a worker's lease: stop waiting for it (Q8/Q9)
                            ex.members_changed.clear()
                            gone = [w for w in workers if w not in ex.members]
                            if gone:
                                self.log(f"[ps {spec.job_id[:8]}] round {rnd}: scheduler removed {gone}")
                            for w in ex.members:           # newcomers are admitted at the next round boundary
                                if w not in workers and w not in pending_join:
                                    pending_join.append(w)
                                    allowed.add(w)
                            workers = [w for w in workers if w in ex.members]
                            result_peers = [w for w in result_peers if w in ex.members or w in pending_join]
                            for w in gone:
                                p_gone = got.pop(w, None)
                                if p_gone and os.path.exists(p_gone):
                                    os.unlink(p_gone)
                            if not workers:
                                raise RuntimeError("no train worker left")
                        if nxt in done:
                            st = nxt.result()
                            nxt = None
                            if st.peer not in workers:
                                st.discard()                           # late file from a removed worker: its sender
                                continue                               # runs out of credit, the link stays usable
                            path = os.path.join(work_dir, hashlib.sha256(st.peer.encode()).hexdigest())
                            try:
                                await st.save(path)
                            except NetworkError as exc:
                                # the worker died while pushing: half a file is no pseudo-gradient.  Keep waiting;
                                # the scheduler will shrink the membership (or the receive timeout drops the worker)
                                self.log(f"[ps {spec.job_id[:8]}] round {rnd}: pseudo-gradient of {st.peer} incomplete ({exc})")
                                if os.path.exists(path):
                                    os.unlink(path)
                                continue
                            got[st.peer] = path
                        elif not done:
                            missing = [w for w in workers if w not in got]
                            if not got:
                                raise asyncio.TimeoutError()
                            self.log(f"[ps {spec.job_id[:8]}] round {rnd}: timeout, dropping {missing}")
                            workers = [w for w in workers if w in got]      # membership shrink (Q8)
                            break
                    # true mean + momentum + Nesterov + cumulative update in ONE pass over the mmapped files
                    # (C++ core, all cores; torch fallback for non-F32 tensors), OFF the event loop: the worker
                    # must keep answering RenewLease / health RPCs while GBs are being reduced
                    # a worker file is untrusted input: one with a malformed header costs that worker its
                    # contribution to this round, not everybody the round
                    for peer, path in list(got.items()):
                        try:
                            read_header(path)
                        except (SafetensorsFormatError, OSError) as exc:
                            self.log(f"[ps {spec.job_id[:8]}] round {rnd}: dropping the file of {peer}: {exc}")
                            got.pop(peer)
                            if os.path.exists(path):
                                os.unlink(path)
                    if not got:
                        raise RuntimeError(f"round {rnd}: no usable pseudo-gradient file")
                    paths = list(got.values())
                    stats = await asyncio.get_event_loop().run_in_executor(
                        None, lambda: outer_round(paths, work_dir, cfg.optimizer.learning_rate, cfg.optimizer.momentum))
                    out = stats["path"]
                    self.log(f"[ps {spec.job_id[:8]}] round {rnd}: reduced {len(paths)} x {stats['elements']} elements in "
                             f"{stats['seconds'] * 1e3:.1f} ms ({'native, ' + str(stats['threads']) + ' threads' if stats['native'] else 'torch fallback'})"
                             + (f", sample weights {[round(w, 3) for w in stats['weights']]}" if stats.get("weights") else ""))
                    # Tell the scheduler the round is complete BEFORE distributing the result: a worker
                    # that gets its file early would otherwise ack ("update-received") against the old
                    # round number and be told to continue (the reference sends results first,
                    # parameter_server.rs:232-282, and has this race).
                    resp = await self.network.request(scheduler, M.PROGRESS_PROTOCOL,
                                                      M.ProgressRequest(spec.job_id, M.Progress("updated")))
                    send_ref = M.Send.to_peers([p for p in result_peers if p in got] or result_peers)
                    # a worker that died since it pushed its pseudo-gradient must not cost the others the round
                    await self.connector.send(send_ref, out, job_id=spec.job_id, tolerate_partial=True)
                    for p, why in self.connector.last_send_failures:
                        self.log(f"[ps {spec.job_id[:8]}] round {rnd}: update not delivered to {p}: {why}")
                    for p in got.values():
                        os.unlink(p)
                    os.unlink(out)
                    rnd += 1
                    if ckpt_dir and rnd % ckpt_every == 0:
                        os.makedirs(ckpt_dir, exist_ok=True)
                        await asyncio.get_event_loop().run_in_executor(
                            None, save_job_checkpoint, ckpt_dir, rnd, cum_path, os.path.join(work_dir, "momentum"))
                    if M.ProgressResponse.from_wire(resp).type == "Done":
                        ex.status = M.JobStatus.FINISHED
                        return
            except asyncio.CancelledError:
                ex.status = M.JobStatus.FAILED
            except Exception as exc:
                self.log(f"[ps {spec.job_id[:8]}] failed: {type(exc).__name__}: {exc}")
                ex.status = M.JobStatus.FAILED
            finally:
                if nxt is not None and not nxt.done():
                    nxt.cancel()
                try:
                    await incoming.aclose()
                except Exception:
                    pass
                shutil.rmtree(work_dir, ignore_errors=True)
                ex.done.set()

        task = asyncio.ensure_future(run())

        async def watch_cancel():
            await ex._cancel.wait()
            task.cancel()

        ex._task = task  # type: ignore[attr-defined]
        ex._watch = asyncio.ensure_future(watch_cancel())  # type: ignore[attr-defined]
        return ex


# ---------------------------------------------------------------------------------------------
# job manager
# ---------------------------------------------------------------------------------------------
@dataclass
class Job:
    id: str
    lease_id: str
    scheduler: str
    spec: M.JobSpec
    execution: Execution


class JobManagerError(Exception):
    pass


class JobManager:
    """Picks the executor for a dispatched job and tracks running jobs by id / lease."""

    def __init__(self, connector: Connector, network: Network, work_dir: str, executor_configs: List[Any],
                 log: Callable[[str], None] = print, executor_env: Optional[Dict[str, str]] = None):
        self.connector, self.network, self.work_dir, self.log = connector, network, work_dir, log
        self.executor_configs = executor_configs
        self.executor_env = dict(executor_env or {})      # extra environment of process executors (HYPHA_DEVICE)
        self.on_job_end: Optional[Callable[[Job, M.JobStatus], None]] = None    # set by the arbiter
        self.jobs: Dict[str, Job] = {}

    def supported(self) -> List[M.ExecutorDescriptor]:
        return [M.ExecutorDescriptor(e.cls, e.name) for e in self.executor_configs]

    def _find(self, desc: M.ExecutorDescriptor):
        for e in self.executor_configs:
            if e.cls == desc.cls and e.name == desc.name:
                return e
        raise JobManagerError(f"no executor configured for {desc.cls}/{desc.name}")

    async def execute(self, task_id: str, spec: M.JobSpec, lease_id: str, scheduler: str) -> Job:
        ec = self._find(spec.executor.descriptor)
        if ec.runtime == "process":
            env = dict(self.executor_env)
            env.update({str(k): str(v) for k, v in (getattr(ec, "env", None) or {}).items()})
            runner = ProcessExecutor(self.connector, self.network, self.work_dir, ec.cmd, list(ec.args),
                                     env=env, log=self.log)
        elif ec.runtime == "parameter-server":
            if spec.executor.descriptor.cls != "aggregate":
                raise JobManagerError("parameter-server runtime only runs aggregate jobs")
            runner = ParameterServerExecutor(self.connector, self.network, self.work_dir, log=self.log)
        else:
            raise JobManagerError(f"unknown runtime {ec.runtime!r}")
        execution = await runner.execute(spec, task_id, scheduler)
        job = Job(task_id, lease_id, scheduler, spec, execution)
        self.jobs[task_id] = job
        asyncio.ensure_future(self._report(job))
        return job

    async def _report(self, job: Job) -> None:
        status = await job.execution.wait()
        if self.on_job_end is not None:
            try:
                self.on_job_end(job, status)
            except Exception as exc:
                self.log(f"[jobs] on_job_end failed: {exc}")
        try:
            await self.network.request(job.scheduler, M.API_PROTOCOL, M.ApiRequest.job_status(job.id, status),
                                       timeout=5.0)
        except NetworkError:
            pass

    def find_jobs_by_lease(self, lease_id: str) -> List[Job]:
        return [j for j in self.jobs.values() if j.lease_id == lease_id]

    def find_jobs_by_scheduler(self, scheduler: str) -> List[Job]:
        return [j for j in self.jobs.values() if j.scheduler == scheduler]

    def get_status(self, task_id: str) -> M.JobStatus:
        j = self.jobs.get(task_id)
        return j.execution.status if j else M.JobStatus.UNKNOWN

    async def cancel(self, task_id: str) -> None:
        j = self.jobs.get(task_id)
        if j is not None:
            j.execution.cancel()
            await j.execution.wait()

    def cleanup_finished_jobs(self) -> int:
        gone = [k for k, j in self.jobs.items() if j.execution.done.is_set()]
        for k in gone:
            del self.jobs[k]
        return len(gone)

    async def shutdown(self) -> None:
        for j in list(self.jobs.values()):
            j.execution.cancel()
        for j in list(self.jobs.values()):
            try:
                await asyncio.wait_for(j.execution.wait(), KILL_GRACE_S + 2)
            except asyncio.TimeoutError:
                pass
