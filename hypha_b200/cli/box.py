"""``hypha-box``: one command for one box (not in the reference, whose roles are always separate
daemons).  Starts gateway + data node + one worker per GPU + scheduler in one process over the
loopback fabric and runs ONE DiLoCo job to completion, in ``collective`` aggregation mode by default:
the scheduler leases the workers, every worker spawns the executor for its GPU, the executors train
through the sm_100a kernels and synchronise with the fused NVLink outer step.

    hypha-box run --model llama3-8b --workers 8 --rounds 10 --samples-per-round 6400 --batch-size 2
    hypha-box run --model-config ./my-llama/config.json --dataset-path ./corpus
    hypha-box run --model gpt2-tiny --workers 2 --device cpu              # plumbing check without GPUs

The job is the same `[scheduler.job]` a stand-alone ``hypha-scheduler`` takes (``--job`` merges extra
keys as JSON); the result is printed as one JSON line.
"""
from __future__ import annotations

import argparse
import asyncio
import json
import os
import shutil
import sys
import tempfile
from typing import List, Optional


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(prog="hypha-box", description=__doc__.split("\n\n")[0])
    sub = p.add_subparsers(dest="command", required=True)
    r = sub.add_parser("run", help="run one DiLoCo job on this box")
    r.add_argument("--model", default="llama-tiny",
                   help="preset (llama3-8b, llama3-1b, llama-small, llama-tiny, gpt2-tiny, gpt2-small)")
    r.add_argument("--model-config", default=None, help="path to a Hugging-Face style config.json (overrides --model)")
    r.add_argument("--model-dir", default=None,
                   help="Hugging Face model directory (config.json + model.safetensors or its shards + tokenizer files): "
                        "training starts from these weights (overrides --model / --model-config)")
    r.add_argument("--workers", type=int, default=0, help="train workers (default: one per visible GPU, 2 on CPU)")
    r.add_argument("--device", choices=["auto", "cuda", "cpu"], default="auto")
    r.add_argument("--rounds", type=int, default=2, help="DiLoCo outer rounds")
    r.add_argument("--samples-per-round", type=int, default=64, help="samples processed by all workers between outer steps")
    r.add_argument("--batch-size", type=int, default=2, help="per-worker batch size cap")
    r.add_argument("--seq-len", type=int, default=128, help="sequence length of the synthetic corpus")
    r.add_argument("--dataset-path", default=None, help="directory of safetensors slices (default: synthetic tokens)")
    r.add_argument("--lr", type=float, default=1e-4)
    r.add_argument("--outer-lr", type=float, default=0.7)
    r.add_argument("--outer-momentum", type=float, default=0.9)
    r.add_argument("--aggregation", choices=["collective", "parameter-server"], default="collective")
    r.add_argument("--precision", choices=["bf16", "mxfp8", "fp8"], default="bf16",
                   help="GEMM precision of the inner steps on GPU executors ([scheduler.job.precision] gemm)")
    r.add_argument("--micro-batches", type=int, default=1,
                   help="gradient accumulation: micro-batches per inner step ([scheduler.job.precision] micro_batches)")
    r.add_argument("--main-grad", choices=["fp32", "bf16"], default="fp32",
                   help="where micro-batch gradients are summed ([scheduler.job.precision] main_grad)")
    r.add_argument("--checkpoint-dir", default=None,
                   help="write a job checkpoint after every --checkpoint-every-th round ([scheduler.job.checkpoint] dir)")
    r.add_argument("--checkpoint-every", type=int, default=1)
    r.add_argument("--resume", action="store_true", help="continue from the newest checkpoint in --checkpoint-dir")
    r.add_argument("--export", default=None, metavar="DIR",
                   help="after the run, write the trained model as a Hugging Face directory (needs --checkpoint-dir)")
    r.add_argument("--job", default=None, help="extra [scheduler.job] keys as JSON (merged last)")
    r.add_argument("--work-dir", default=None, help="scratch directory (default: a temporary one, removed at exit)")
    r.add_argument("--timeout", type=float, default=3600.0)
    r.add_argument("--verbose", action="store_true", help="echo executor output")
    e = sub.add_parser("export", help="turn a job checkpoint into a model directory `transformers` loads",
                       description="Rebuild the global model of a job checkpoint ([scheduler.job.checkpoint] dir) -- "
                                   "theta_0 + cumulative update of a parameter-server job, theta_prev of a collective "
                                   "job, flat parameters + fp32 shards of the fused GPU engine -- and write "
                                   "model.safetensors under Hugging Face names next to config.json.")
    e.add_argument("--checkpoint", required=True, help="checkpoint directory of the job")
    e.add_argument("--model", required=True, help="model directory the job trained (config.json [+ model.safetensors])")
    e.add_argument("-o", "--output", required=True)
    e.add_argument("--round", type=int, default=None, help="round to export (default: the newest)")
    e.add_argument("--dtype", choices=["bf16", "fp16", "fp32"], default=None)
    e.add_argument("--type", default="causal-lm", help="model type of the job ([scheduler.job.model] type)")
    v = sub.add_parser("eval", help="next-token loss / perplexity of a model directory over a slice directory",
                       description="Run the model of --model (a Hugging Face directory, e.g. the output of `export`) "
                                   "over the slices of --dataset-path with the executors' model code and print "
                                   "loss and perplexity as one JSON line.")
    v.add_argument("--model", required=True)
    v.add_argument("--dataset-path", required=True)
    v.add_argument("--device", choices=["auto", "cuda", "cpu"], default="auto")
    v.add_argument("--batch-size", type=int, default=4)
    v.add_argument("--max-samples", type=int, default=None)
    v.add_argument("--type", default="causal-lm")
    return p


MODEL_DIR_FILES = ("model.safetensors", "model.safetensors.index.json", "generation_config.json", "tokenizer.json",
                   "tokenizer_config.json", "tokenizer.model", "special_tokens_map.json", "vocab.json", "merges.txt")


def model_dir_files(model_dir: str) -> List[str]:
    """What a job over ``model_dir`` lists as its model `filenames`: the config, the weights (one file or every shard of
    the index) and the tokenizer files that are there."""
    if not os.path.isfile(os.path.join(model_dir, "config.json")):
        raise FileNotFoundError(f"{model_dir} has no config.json")
    names = ["config.json"] + [n for n in MODEL_DIR_FILES if os.path.isfile(os.path.join(model_dir, n))]
    index = os.path.join(model_dir, "model.safetensors.index.json")
    if os.path.isfile(index):
        names += sorted(set(json.load(open(index)).get("weight_map", {}).values()))
    return names


def _write_model_config(args) -> tuple:
    """Returns (repo name, writer(path), vocab_size)."""
    if getattr(args, "model_dir", None):
        src = os.path.abspath(args.model_dir)
        meta = json.load(open(os.path.join(src, "config.json")))

        def link_all(cfg_path: str) -> None:
            # the registry entry of the job's model: links, not copies (an 8B checkpoint is 16 GB)
            dst = os.path.dirname(cfg_path)
            for n in model_dir_files(src):
                target = os.path.join(dst, n)
                if os.path.lexists(target):
                    os.unlink(target)
                os.symlink(os.path.join(src, n), target)

        return f"local/{os.path.basename(src.rstrip('/')) or 'model'}", link_all, int(meta.get("vocab_size", 32000))
    if args.model_config:
        meta = json.load(open(args.model_config))
        return "local/custom", (lambda p: shutil.copyfile(args.model_config, p)), int(meta.get("vocab_size", 32000))
    if args.model.startswith("gpt2"):
        from ..models.gpt2 import GPT2Config

        cfg = GPT2Config.tiny() if args.model == "gpt2-tiny" else GPT2Config()
        return f"local/{args.model}", cfg.to_json, int(cfg.vocab_size)
    from ..models.llama import get_config

    cfg = get_config(args.model, max_seq_len=max(args.seq_len, 8))
    return f"local/{args.model}", (lambda p: cfg.to_hf_json(p)), int(cfg.vocab_size)


def _deep_merge(a: dict, b: dict) -> dict:
    for k, v in b.items():
        a[k] = _deep_merge(a.get(k, {}), v) if isinstance(v, dict) and isinstance(a.get(k), dict) else v
    return a


def run(args) -> int:
    import torch

    from ..executor.dataset import write_synthetic_token_slices
    from ..local_cluster import run_loopback_job

    use_cuda = args.device == "cuda" or (args.device == "auto" and torch.cuda.is_available())
    workers = args.workers or (torch.cuda.device_count() if use_cuda else 2)
    if use_cuda and workers > torch.cuda.device_count():
        print(f"hypha-box: {workers} workers requested but only {torch.cuda.device_count()} GPUs visible", file=sys.stderr)
        return 2
    os.environ["HYPHA_EXECUTOR_DEVICE"] = "cuda" if use_cuda else "cpu"
    repo, writer, vocab = _write_model_config(args)
    job = {
        "model": {"repository": repo, "filenames": model_dir_files(args.model_dir) if args.model_dir else ["config.json"],
                  "type": "causal-lm", "input_names": ["input_ids", "labels"]},
        "preprocessor": None,
        "dataset": {"dataset": os.path.basename(os.path.normpath(args.dataset_path)) if args.dataset_path else "synthetic-tokens"},
        "rounds": {"avg_samples_between_updates": args.samples_per_round, "update_rounds": args.rounds,
                   "max_batch_size": args.batch_size},
        "inner_optimizer": {"learning-rate": args.lr},
        "outer_optimizer": {"learning-rate": args.outer_lr, "momentum": args.outer_momentum},
        "aggregation": {"mode": args.aggregation},
        "precision": {"gemm": args.precision, "micro_batches": args.micro_batches, "main_grad": args.main_grad},
        "resources": {"num_workers": workers, "worker": {"gpu": 1.0, "cpu": 1.0, "memory": 1.0, "storage": 0.0},
                      "parameter_server": {"gpu": 0.0, "cpu": 1.0, "memory": 1.0, "storage": 0.0},
                      "worker_price": {"bid": 1.0, "max": 2.0}, "parameter_server_price": {"bid": 1.0, "max": 2.0}},
    }
    if args.export and not args.checkpoint_dir:
        print("hypha-box: --export needs --checkpoint-dir (the model is rebuilt from the job checkpoint)", file=sys.stderr)
        return 2
    if args.checkpoint_dir:
        os.makedirs(args.checkpoint_dir, exist_ok=True)
        job["checkpoint"] = {"dir": os.path.abspath(args.checkpoint_dir), "every": args.checkpoint_every,
                             "resume": bool(args.resume)}
    if args.job:
        _deep_merge(job, json.loads(args.job))
    tmp = args.work_dir or tempfile.mkdtemp(prefix="hypha-box-")
    os.makedirs(tmp, exist_ok=True)

    def write_dataset(d: str) -> None:
        if args.dataset_path:
            if os.path.abspath(args.dataset_path) != os.path.abspath(d):
                shutil.copytree(args.dataset_path, d, dirs_exist_ok=True)
            return
        per_round = max(args.samples_per_round, workers * args.batch_size)
        slices = max(2 * workers, 4)
        write_synthetic_token_slices(d, num_slices=slices, samples_per_slice=max(8, -(-per_round * 2 // slices)),
                                     seq_len=args.seq_len, vocab_size=vocab)

    log = (lambda m: print(m, file=sys.stderr, flush=True)) if args.verbose else None
    exported = None
    try:
        outcome, sink, logs, _ = asyncio.run(run_loopback_job(
            tmp, repo=repo, write_model_config=writer, write_dataset=write_dataset, job=job, num_workers=workers,
            timeout_s=args.timeout, log=log))
        if args.export and outcome.finished:
            from ..export import ExportError, export_checkpoint

            mdir = os.path.join(tmp, "export-model")
            os.makedirs(mdir, exist_ok=True)
            writer(os.path.join(mdir, "config.json"))        # --model-dir: links to the weights the job started from
            try:
                exported = export_checkpoint(args.checkpoint_dir, mdir, args.export)
            except ExportError as exc:
                print(f"hypha-box: export failed: {exc}", file=sys.stderr)
    finally:
        if not args.work_dir:
            shutil.rmtree(tmp, ignore_errors=True)
    by_worker: dict = {}
    for rec in sink.records:
        by_worker.setdefault(rec["worker_id"], {}).setdefault(str(rec["round"]), {})[rec["metric_name"]] = rec["value"]
    print(json.dumps({"finished": outcome.finished, "rounds": outcome.rounds, "workers": outcome.workers,
                      "removed_workers": outcome.removed_workers, "aggregation": args.aggregation, "precision": args.precision,
                      "device": "cuda" if use_cuda else "cpu", "metrics": by_worker,
                      **({"exported": exported} if exported else {})}))
    if not outcome.finished:
        print("\n".join(logs[-30:]), file=sys.stderr)
    return 0 if outcome.finished else 1


def main(argv: Optional[List[str]] = None) -> int:
    args = build_parser().parse_args(argv)
    if args.command == "export":
        from ..export import ExportError, export_checkpoint

        try:
            info = export_checkpoint(args.checkpoint, args.model, args.output, model_type=args.type, round_no=args.round,
                                     dtype=args.dtype)
        except (ExportError, OSError) as exc:
            print(f"hypha-box export: {exc}", file=sys.stderr)
            return 2
        print(json.dumps(info))
        return 0
    if args.command == "eval":
        import torch

        from ..export import ExportError, evaluate_model

        dev = "cuda" if args.device == "cuda" or (args.device == "auto" and torch.cuda.is_available()) else "cpu"
        try:
            print(json.dumps(evaluate_model(args.model, args.dataset_path, model_type=args.type, device=dev,
                                            batch_size=args.batch_size, max_samples=args.max_samples)))
        except (ExportError, OSError) as exc:
            print(f"hypha-box eval: {exc}", file=sys.stderr)
            return 2
        return 0
    return run(args)


if __name__ == "__main__":
    sys.exit(main())
