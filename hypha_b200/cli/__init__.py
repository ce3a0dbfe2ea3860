"""Role command lines: ``hypha-{gateway,scheduler,worker,data} {init,probe,run}`` and
``hypha-certutil {root,org,node}``.

Reference: crates/*/src/cli.rs (``init`` writes a documented TOML, ``run -c file`` layers
TOML -> HYPHA_* -> OTEL_* -> flags, ``probe <addr>`` exits 0 iff the node reports healthy;
the scheduler's ``init --job '<json>'`` injects a job: hypha-scheduler.rs:477-545)."""
from __future__ import annotations

import argparse
import asyncio
import json
import signal
import sys
from typing import Any, Dict, List, Optional

from .. import config as C
from .. import roles

ROLES = {
    "gateway": (C.GatewayConfig, roles.run_gateway),
    "scheduler": (C.SchedulerConfig, roles.run_scheduler),
    "worker": (C.WorkerConfig, roles.run_worker),
    "data": (C.DataConfig, roles.run_data),
}


def build_parser(role: str) -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(prog=f"hypha-{role}", description=(ROLES[role][0].__doc__ or "").strip())
    sub = p.add_subparsers(dest="command", required=True)
    i = sub.add_parser("init", help="write a documented default configuration")
    i.add_argument("-o", "--output", default=f"{role}.toml")
    i.add_argument("-n", "--name", default=None, help="node name (used for certificate file names)")
    if role == "scheduler":
        i.add_argument("--job", default=None, help="job definition as JSON, merged into [scheduler.job]")
    if role == "data":
        i.add_argument("--dataset-path", default=None)
    if role == "worker":
        i.add_argument("--detect", action="store_true",
                       help="fill [resources] from this machine (visible GPUs, cores, RAM, free disk)")
        i.add_argument("--device", type=int, default=None, help="CUDA device index of this worker's executors")
    pr = sub.add_parser("probe", help="check a node's health endpoint")
    pr.add_argument("address")
    pr.add_argument("--timeout", type=float, default=5.0)
    r = sub.add_parser("run", help="run the role")
    r.add_argument("-c", "--config", default=None, help="TOML configuration file")
    r.add_argument("--gateway", dest="gateway_addresses", action="append", default=None)
    r.add_argument("--listen", dest="listen_addresses", action="append", default=None)
    r.add_argument("--external", dest="external_addresses", action="append", default=None)
    r.add_argument("--exclude-cidr", dest="exclude_cidr", action="append", default=None)
    r.add_argument("--peer-id", default=None)
    if role == "worker":
        r.add_argument("--work-dir", default=None)
        r.add_argument("--device", type=int, default=None, help="CUDA device index of this worker's executors")
    if role == "scheduler":
        r.add_argument("--status-bridge", default=None)
    if role == "data":
        r.add_argument("--dataset-path", default=None)
    if role == "data":
        pp = sub.add_parser("prepare", help="cut a token corpus into the slice files a data node serves",
                            description="Pack a pre-tokenised corpus (.npy | .bin | .safetensors) or text (.txt | "
                                        ".jsonl with a local tokenizer) into samples of --seq-len ids and write "
                                        "them as safetensors slices of --samples-per-slice rows into --output; "
                                        "point `dataset_path` of a data node at that directory.")
        pp.add_argument("source", help="corpus file")
        pp.add_argument("-o", "--output", required=True, help="slice directory to create (its name = dataset name)")
        pp.add_argument("--seq-len", type=int, required=True)
        pp.add_argument("--samples-per-slice", type=int, default=1024)
        pp.add_argument("--dtype", default="uint16", help="id type of a .bin corpus (uint16 | uint32 | int32 | int64)")
        pp.add_argument("--tokenizer", default=None, help="local directory of a Hugging Face tokenizer (text input)")
        pp.add_argument("--vocab-size", type=int, default=None, help="refuse ids outside 0..vocab-1")
        pp.add_argument("--shuffle-seed", type=int, default=None, help="permute the samples before slicing")
        pp.add_argument("--no-labels", action="store_true", help="write input_ids only")
        pp.add_argument("--drop-last", action="store_true", help="drop a final slice shorter than the others")
    return p


def cli_markdown(role: str) -> str:
    """CLI reference as markdown (the reference renders this in build.rs: crates/*/build.rs:10-28)."""
    p = build_parser(role)
    out = [f"# hypha-{role}", "", "```", p.format_help().rstrip(), "```", ""]
    for name, sp in p._subparsers._group_actions[0].choices.items():  # type: ignore[union-attr]
        out += [f"## hypha-{role} {name}", "", "```", sp.format_help().rstrip(), "```", ""]
    return "\n".join(out)


def detect_resources(work_dir: str = "/tmp", one_gpu: bool = False) -> Dict[str, float]:
    """What `hypha-worker init --detect` writes into [resources]: GPUs visible to CUDA (1 when the
    worker is pinned to one device), CPU cores, RAM in GB, free disk under ``work_dir`` in GB."""
    import os
    import shutil

    gpus = 0
    try:
        import torch

        gpus = torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        gpus = 0
    if one_gpu and gpus > 0:
        gpus = 1
    try:
        import psutil

        mem_gb = psutil.virtual_memory().total / 2 ** 30
    except Exception:
        mem_gb = os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES") / 2 ** 30
    disk_gb = shutil.disk_usage(work_dir if os.path.isdir(work_dir) else "/").free / 2 ** 30
    return {"gpu": float(gpus), "cpu": float(os.cpu_count() or 1), "memory": float(int(mem_gb)),
            "storage": float(int(disk_gb))}


def _overrides(ns: argparse.Namespace, keys: List[str]) -> Dict[str, Any]:
    return {k: getattr(ns, k) for k in keys if getattr(ns, k, None) is not None}


def main(role: str, argv: Optional[List[str]] = None) -> int:
    cls, runner = ROLES[role]
    ns = build_parser(role).parse_args(argv)
    if ns.command == "init":
        over: Dict[str, Any] = {}
        if ns.name:
            over.update(cert_pem=f"{ns.name}-cert.pem", key_pem=f"{ns.name}-key.pem", trust_pem=f"{ns.name}-trust.pem")
        if role == "scheduler" and ns.job:
            over["scheduler"] = {"job": json.loads(ns.job)}
        if role == "data" and ns.dataset_path:
            over["dataset_path"] = ns.dataset_path
        if role == "worker":
            if ns.detect:
                over["resources"] = detect_resources(one_gpu=ns.device is not None)
            if ns.device is not None:
                over["device"] = ns.device
        cfg = C.builder(cls).with_overrides(over).build()
        with open(ns.output, "w") as f:
            f.write(C.to_toml(cfg))
        print(f"wrote {ns.output}")
        return 0
    if ns.command == "prepare":
        from ..executor.dataset import load_token_corpus, write_token_slices

        try:
            tokens = load_token_corpus(ns.source, dtype=ns.dtype, tokenizer=ns.tokenizer)
            summary = write_token_slices(tokens, ns.output, ns.seq_len, ns.samples_per_slice,
                                         with_labels=not ns.no_labels, vocab_size=ns.vocab_size,
                                         shuffle_seed=ns.shuffle_seed, drop_last=ns.drop_last)
        except (ValueError, OSError) as exc:
            print(f"hypha-data prepare: {exc}", file=sys.stderr)
            return 2
        print(json.dumps(summary))
        return 0
    if ns.command == "probe":
        ok = asyncio.run(roles.probe(ns.address, ns.timeout))
        print("healthy" if ok else "unhealthy")
        return 0 if ok else 1
    keys = ["gateway_addresses", "listen_addresses", "external_addresses", "exclude_cidr", "work_dir",
            "status_bridge", "dataset_path", "device"]
    try:
        cfg = C.validate(C.builder(cls).with_toml(ns.config).with_env().with_overrides(_overrides(ns, keys)).build())
    except C.ConfigError as exc:
        print(f"configuration error: {exc}", file=sys.stderr)
        return 2

    async def amain() -> int:
        loop = asyncio.get_event_loop()
        started = loop.create_future()
        task = asyncio.ensure_future(runner(cfg, peer_id=ns.peer_id, started=started))

        def stop(*_):
            if started.done() and not started.cancelled() and started.exception() is None:
                started.result().stop.set()
            else:
                task.cancel()

        for sig in (signal.SIGINT, signal.SIGTERM):
            loop.add_signal_handler(sig, stop)
        try:
            out = await task
        except asyncio.CancelledError:
            return 130
        if role == "scheduler":
            print(json.dumps({"job_id": out.job_id, "rounds": out.rounds, "finished": out.finished,
                              "workers": out.workers, "parameter_server": out.parameter_server,
                              "removed_workers": out.removed_workers, "joined_workers": out.joined_workers}))
            return 0 if out.finished else 1
        return 0

    try:
        return asyncio.run(amain())
    except roles.RoleError as exc:
        print(f"error: {exc}", file=sys.stderr)
        return 1
