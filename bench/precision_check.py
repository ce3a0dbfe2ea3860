#!/usr/bin/env python
"""Precision evidence for the optimizer-state choice of the headline ("bf16 moments + stochastic
rounding is a precision reduction in optimizer state").

Trains `llama-small` (4 layers, hidden 1024, head_dim 128 -> the hand-written attention) for N steps on a
LEARNABLE synthetic stream (a noisy affine Markov chain over the vocabulary, so the loss really falls) three
times from the same initial weights and data order:

  fp32      fused flat AdamW, fp32 moments                         (DiLoCoTrainer, --state-dtype fp32)
  bf16sr    fused flat AdamW, bf16 moments + stochastic rounding   (the bench default)
  torch     stock torch.optim.AdamW (fp32 master copy of the flat buffer, fp32 moments) driven by the same
            forward/backward -- the reference's optimizer (executors/accelerate/.../utils.py:56-64)

and reports the loss curves (EMA-smoothed) and their relative distance.  -> gpurun_out/precision_check.json

``--kinds bf16sr,mxfp8,fp8tensor`` compares the fp8 GEMM paths against the bf16 run instead (same optimizer):
  mxfp8      every Linear GEMM (fwd, dgrad, wgrad) on tcgen05.mma.kind::mxf8f6f4.block_scale (e4m3 + UE8M0 / 32)
  fp8tensor  per-tensor current scaling on kind::f8f6f4 (e4m3 forward, e5m2 gradients)"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypha_b200.models.llama import get_config  # noqa: E402
from hypha_b200.parallel.diloco import AdamWConfig, DiLoCoConfig  # noqa: E402
from hypha_b200.trainer import DiLoCoTrainer  # noqa: E402


def make_stream(vocab, batch, seq, steps, seed=0):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(steps):
        x = torch.empty(batch, seq, dtype=torch.int64)
        x[:, 0] = torch.randint(0, vocab, (batch,), generator=g)
        noise = torch.rand(batch, seq, generator=g) < 0.1
        rnd = torch.randint(0, vocab, (batch, seq), generator=g)
        for t in range(1, seq):
            nxt = (x[:, t - 1] * 31 + 7) % vocab
            x[:, t] = torch.where(noise[:, t], rnd[:, t], nxt)
        out.append(x)
    return out


def run(kind, cfg, stream, lr, init):
    from hypha_b200 import ops

    ops.set_fp8(kind in ("mxfp8", "fp8tensor"), mode="mx" if kind == "mxfp8" else "tensor")
    tr = DiLoCoTrainer.create(cfg, adamw=AdamWConfig(lr=lr, state_dtype="fp32" if kind in ("fp32", "torch") else "bf16"),
                              diloco=DiLoCoConfig(inner_steps=1 << 60), device=0, seed=11)
    if init is not None:
        tr.model.flat.data.copy_(init)
        tr.opt.master.copy_(init)
    else:
        init = tr.model.flat.data.clone()
    losses = []
    if kind == "torch":
        tr.model.grad_ready_hook = None        # no fused AdamW slices during backward: torch's optimizer steps instead
        master = torch.nn.Parameter(tr.model.flat.data.float().clone())
        opt = torch.optim.AdamW([master], lr=lr)                   # torch defaults: betas (0.9, 0.999), wd 0.01
        for ids in stream:
            ids = ids.cuda()
            tr.model.set_grad_accumulate(False)
            loss = tr.model(ids, torch.roll(ids, -1, 1))
            loss.backward()
            master.grad = tr.model.flat.grad.float()
            opt.step()
            tr.model.flat.data.copy_(master.data)
            losses.append(float(loss.item()))
    else:
        for ids in stream:
            losses.append(float(tr.step(ids.pin_memory()).item()))
    del tr
    torch.cuda.empty_cache()
    ops.set_fp8(False)
    return losses, init


def ema(xs, a=0.9):
    out, m = [], xs[0]
    for x in xs:
        m = a * m + (1 - a) * x
        out.append(m)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--seq", type=int, default=512)
    ap.add_argument("--lr", type=float, default=3e-4)
    ap.add_argument("--kinds", default="fp32,bf16sr,torch")
    ap.add_argument("--out", default="gpurun_out/precision_check.json")
    a = ap.parse_args()
    kinds = a.kinds.split(",")
    cfg = get_config("llama-small", max_seq_len=a.seq)
    stream = make_stream(cfg.vocab_size, a.batch, a.seq, a.steps)
    curves, init = {}, None
    for kind in kinds:
        curves[kind], init = run(kind, cfg, stream, a.lr, init)
        print(kind, [round(curves[kind][i], 4) for i in (0, a.steps // 4, a.steps // 2, a.steps - 1)], flush=True)
    sm = {k: ema(v) for k, v in curves.items()}
    tail = slice(a.steps // 3, None)
    pairs = [(x, y) for x, y in (("bf16sr", "fp32"), ("fp32", "torch"), ("bf16sr", "torch"), ("mxfp8", "bf16sr"),
                                 ("fp8tensor", "bf16sr")) if x in sm and y in sm]
    rel = {f"{x}_vs_{y}": max(abs(p - q) / q for p, q in zip(sm[x][tail], sm[y][tail])) for x, y in pairs}
    out = {"model": "llama-small", "steps": a.steps, "batch": a.batch, "seq": a.seq, "lr": a.lr,
           "loss_at": {k: {str(i): v[i] for i in (0, 49, 99, 199, a.steps - 1) if i < a.steps} for k, v in sm.items()},
           "max_rel_diff_of_smoothed_loss_after_first_third": rel, "curves": curves}
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(a.out, "w"))
    print(json.dumps({k: v for k, v in out.items() if k != "curves"}))


if __name__ == "__main__":
    main()
