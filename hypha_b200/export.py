"""Turn a job checkpoint into a model directory `transformers` can load.

The reference keeps nothing of a finished job: the global model only exists as transient files of the workers and the
parameter server, removed at job end (training.py:60,135, parameter_server.rs:222-427).  hypha_b200 writes job
checkpoints (`[scheduler.job.checkpoint]`); this module is the way back out: it rebuilds the global model theta from
whichever of the three checkpoint kinds it finds and writes ``model.safetensors`` under Hugging Face parameter names
next to the model's ``config.json`` (+ tokenizer files, if the model directory has them).

* parameter-server jobs store the cumulative outer update ``U``: theta = theta_0 + U, theta_0 = the model directory's
  weights, or the seeded initialisation every worker started from when it has none;
* collective jobs on torch engines (CPU / other families) store theta_prev per tensor (``collective_state``);
* collective jobs on the fused GPU engine store the flat bf16 parameter vector plus one fp32 theta_outer shard per
  rank -- the shards, when all are there, give the fp32 model.
"""
from __future__ import annotations

import json
import os
import shutil
from typing import Dict, Optional

import torch

_COPY_ALONG = ("config.json", "generation_config.json", "tokenizer.json", "tokenizer_config.json", "tokenizer.model",
               "special_tokens_map.json", "vocab.json", "merges.txt", "preprocessor_config.json")


class ExportError(RuntimeError):
    pass


def _round_dir(ckpt_dir: str, round_no: Optional[int]) -> tuple:
    if round_no is None:
        try:
            round_no = int(json.load(open(os.path.join(ckpt_dir, "latest.json")))["round"])
        except (OSError, ValueError, KeyError) as exc:
            raise ExportError(f"{ckpt_dir} holds no complete checkpoint (latest.json: {exc})") from exc
    path = os.path.join(ckpt_dir, f"round-{round_no}")
    if not os.path.isdir(path):
        raise ExportError(f"no checkpoint of round {round_no} in {ckpt_dir}")
    return round_no, path


def _flat_from_fused(path: str) -> torch.Tensor:
    """The whole model as one fp32 vector from a fused-engine checkpoint: the fp32 theta_outer shards where they are
    all present and cover the vector, else the bf16 parameters."""
    from safetensors.torch import load_file

    flat = load_file(os.path.join(path, "collective_params"))["params"].float()
    shards = sorted(f for f in os.listdir(path) if f.startswith("outer_shard_") and not f.endswith(".tmp"))
    pieces, world = [], None
    for f in shards:
        sh = load_file(os.path.join(path, f))
        _, begin, length, w = (int(x) for x in sh["meta"].tolist())
        world = w if world is None else world
        pieces.append((begin, length, sh["theta_outer"].float()))
    pieces.sort()
    covered, ok = 0, world is not None and len(pieces) == world
    for begin, length, t in pieces:
        if begin != covered or t.numel() != length:
            ok = False
            break
        covered += length
    if ok and covered == flat.numel():
        for begin, length, t in pieces:
            flat[begin:begin + length] = t
    return flat


def _apply(model: torch.nn.Module, tensors: Dict[str, torch.Tensor], add: bool) -> int:
    """theta += tensors (``add``) or theta = tensors, matching either the torch state-dict names of the model or --
    for the Llama family -- the names of the flat sm_100a layout (what GPU executors report).  Returns how many
    tensors were used; raises when nothing matches."""
    sd = model.state_dict()
    used = 0
    with torch.no_grad():
        for k, v in tensors.items():
            if k in sd and sd[k].shape == v.shape:
                sd[k].copy_((sd[k].float() + v.float()) if add else v.float())
                used += 1
        net = getattr(model, "net", None)
        if used == 0 and net is not None and hasattr(net, "flat"):
            for k, v in tensors.items():
                if k in net.flat.layout:
                    dst = net.flat.view(k)
                    dst.copy_((dst.float() + v.float().view(dst.shape)) if add else v.float().view(dst.shape))
                    used += 1
    if used == 0:
        raise ExportError("the checkpoint's tensor names match neither the model's state dict nor its flat layout "
                          f"(checkpoint: {sorted(tensors)[:3]}..., model: {sorted(sd)[:3]}...)")
    return used


def export_checkpoint(ckpt_dir: str, model_dir: str, out_dir: str, model_type: str = "causal-lm",
                      round_no: Optional[int] = None, dtype: Optional[str] = None) -> dict:
    """Write ``out_dir/model.safetensors`` (+ config / tokenizer files) for the global model of checkpoint
    ``round_no`` (default: the newest) of the job that trained ``model_dir``.  ``dtype``: ``bf16`` | ``fp16`` |
    ``fp32`` (default: as stored -- fp32 for everything but a fused checkpoint without its shards)."""
    from safetensors.torch import load_file, save_file

    from .executor.model import get_model

    round_no, path = _round_dir(ckpt_dir, round_no)
    model = get_model(model_dir, model_type, "cpu")             # theta_0: the directory's weights or the seeded init
    model.float()
    have = set(os.listdir(path))
    if {"cumulative_update", "momentum"} <= have:
        kind = "parameter-server"
        used = _apply(model, load_file(os.path.join(path, "cumulative_update")), add=True)
    elif "collective_state" in have:
        kind = "collective"
        st = load_file(os.path.join(path, "collective_state"))
        used = _apply(model, {k.split("/", 1)[1]: v for k, v in st.items() if k.split("/", 1)[0] in ("prev", "buf")},
                      add=False)
    elif "collective_params" in have:
        kind = "collective-fused"
        net = getattr(model, "net", None)
        if net is None or not hasattr(net, "flat"):
            raise ExportError("a fused-engine checkpoint belongs to a model of the Llama family")
        flat = _flat_from_fused(path)
        if flat.numel() != net.flat.data.numel():
            raise ExportError(f"the checkpoint holds {flat.numel()} parameters, the model in {model_dir} has "
                              f"{net.flat.data.numel()}")
        with torch.no_grad():
            net.flat.data.copy_(flat)
        used = len(net.flat.layout)
    else:
        raise ExportError(f"{path} is not a job checkpoint (found {sorted(have)})")
    net = getattr(model, "net", None)
    sd = net.to_hf_state_dict() if net is not None and hasattr(net, "to_hf_state_dict") else model.state_dict()
    cast = {None: None, "fp32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16}
    if dtype not in cast:
        raise ExportError(f"unknown dtype {dtype!r} (bf16 | fp16 | fp32)")
    out, stored = {}, set()
    for k, v in sd.items():
        v = v.detach()
        where = (v.data_ptr(), tuple(v.shape))
        if v.numel() and where in stored:
            continue                    # tied weights (GPT-2's lm_head = wte) go in once; `transformers` re-ties them
        stored.add(where)
        if v.dtype.is_floating_point and cast[dtype] is not None:
            v = v.to(cast[dtype])
        out[k] = v.clone().contiguous()
    os.makedirs(out_dir, exist_ok=True)
    save_file(out, os.path.join(out_dir, "model.safetensors"), metadata={"format": "pt"})
    for name in _COPY_ALONG:
        src = os.path.join(model_dir, name)
        if os.path.isfile(src):
            shutil.copyfile(src, os.path.join(out_dir, name))
    info = {"round": round_no, "kind": kind, "tensors": len(out), "checkpoint_tensors_used": used,
            "parameters": int(sum(v.numel() for v in out.values())), "dtype": dtype or "as stored",
            "output": os.path.join(out_dir, "model.safetensors")}
    return info


@torch.no_grad()
def evaluate_model(model_dir: str, dataset_path: str, model_type: str = "causal-lm", device: str = "cpu",
                   batch_size: int = 4, max_samples: Optional[int] = None) -> dict:
    """Mean next-token loss and perplexity of the model in ``model_dir`` over the slices of ``dataset_path`` (the files a
    data node serves), through the same model code the executors train with -- the sm_100a kernels on ``cuda``.  The loss
    is the Hugging Face one (labels aligned with the inputs, shifted inside, -100 ignored), weighted by the number of
    scored tokens of each batch."""
    from .data import dataset_files
    from .executor.dataset import IterableStreamDataSet
    from .executor.model import get_model

    files = dataset_files(dataset_path)
    if not files:
        raise ExportError(f"{dataset_path} holds no slices")
    model = get_model(model_dir, model_type, device).eval()
    loader = torch.utils.data.DataLoader(IterableStreamDataSet(iter(files), ["input_ids", "labels"]),
                                         batch_size=batch_size)
    total, scored, samples = 0.0, 0, 0
    for batch in loader:
        if max_samples is not None and samples >= max_samples:
            break
        if max_samples is not None:
            batch = {k: v[: max_samples - samples] for k, v in batch.items()}
        n = int((batch["labels"][:, 1:] != -100).sum())
        if n == 0:
            continue
        out = model(**{k: v.to(device) for k, v in batch.items()})
        loss = out["loss"] if isinstance(out, dict) else out.loss
        total += float(loss) * n
        scored += n
        samples += int(batch["input_ids"].shape[0])
    if scored == 0:
        raise ExportError("no scored tokens in the dataset")
    import math

    mean = total / scored
    return {"loss": mean, "perplexity": math.exp(min(mean, 50.0)), "tokens_scored": scored, "samples": samples,
            "device": device}
