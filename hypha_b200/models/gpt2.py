"""GPT-2 family (stock torch ops): the model of the CPU plumbing configuration
(BASELINE config 1: GPT-2 small DiLoCo, 2 workers on CPU via the loopback control plane).

Parameter names and the ``Conv1D`` weight orientation follow HuggingFace ``GPT2LMHeadModel`` so
an HF ``model.safetensors`` + ``config.json`` loads directly (the reference loads every model
through ``AutoModelFor*.from_pretrained``, model.py:47-126)."""
from __future__ import annotations

import json
import math
from dataclasses import asdict, dataclass
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class GPT2Config:
    vocab_size: int = 50257
    n_positions: int = 1024
    n_embd: int = 768
    n_layer: int = 12
    n_head: int = 12
    layer_norm_epsilon: float = 1e-5
    initializer_range: float = 0.02
    model_type: str = "gpt2"

    @classmethod
    def small(cls) -> "GPT2Config":
        return cls()

    @classmethod
    def tiny(cls) -> "GPT2Config":
        return cls(vocab_size=512, n_positions=128, n_embd=64, n_layer=2, n_head=4)

    @classmethod
    def from_json(cls, path: str) -> "GPT2Config":
        d = json.load(open(path))
        keys = cls.__dataclass_fields__.keys()
        return cls(**{k: v for k, v in d.items() if k in keys})

    def to_json(self, path: str) -> None:
        d = asdict(self)
        d["architectures"] = ["GPT2LMHeadModel"]
        json.dump(d, open(path, "w"), indent=1)

    def num_params(self) -> int:
        e, l = self.n_embd, self.n_layer
        return self.vocab_size * e + self.n_positions * e + l * (12 * e * e + 13 * e) + 2 * e


class Conv1D(nn.Module):
    """HF-style linear with weight stored ``[in, out]``."""

    def __init__(self, nx: int, nf: int, std: float):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(nx, nf).normal_(0.0, std))
        self.bias = nn.Parameter(torch.zeros(nf))

    def forward(self, x):
        return torch.addmm(self.bias, x.reshape(-1, x.shape[-1]), self.weight).view(*x.shape[:-1], -1)


class Block(nn.Module):
    def __init__(self, c: GPT2Config):
        super().__init__()
        e = c.n_embd
        self.n_head = c.n_head
        self.ln_1 = nn.LayerNorm(e, eps=c.layer_norm_epsilon)
        self.attn = nn.ModuleDict({"c_attn": Conv1D(e, 3 * e, c.initializer_range),
                                   "c_proj": Conv1D(e, e, c.initializer_range / math.sqrt(2 * c.n_layer))})
        self.ln_2 = nn.LayerNorm(e, eps=c.layer_norm_epsilon)
        self.mlp = nn.ModuleDict({"c_fc": Conv1D(e, 4 * e, c.initializer_range),
                                  "c_proj": Conv1D(4 * e, e, c.initializer_range / math.sqrt(2 * c.n_layer))})

    def forward(self, x):
        B, S, E = x.shape
        q, k, v = self.attn["c_attn"](self.ln_1(x)).split(E, dim=2)
        hd = E // self.n_head
        q, k, v = (t.view(B, S, self.n_head, hd).transpose(1, 2) for t in (q, k, v))
        a = F.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(B, S, E)
        x = x + self.attn["c_proj"](a)
        h = F.gelu(self.mlp["c_fc"](self.ln_2(x)), approximate="tanh")
        return x + self.mlp["c_proj"](h)


class GPT2LMHeadModel(nn.Module):
    def __init__(self, config: GPT2Config):
        super().__init__()
        self.config = config
        c = config
        self.transformer = nn.ModuleDict({
            "wte": nn.Embedding(c.vocab_size, c.n_embd), "wpe": nn.Embedding(c.n_positions, c.n_embd),
            "h": nn.ModuleList([Block(c) for _ in range(c.n_layer)]),
            "ln_f": nn.LayerNorm(c.n_embd, eps=c.layer_norm_epsilon)})
        nn.init.normal_(self.transformer["wte"].weight, std=c.initializer_range)
        nn.init.normal_(self.transformer["wpe"].weight, std=c.initializer_range)

    def forward(self, input_ids: torch.Tensor, labels: Optional[torch.Tensor] = None, **_):
        B, S = input_ids.shape
        pos = torch.arange(S, device=input_ids.device)
        x = self.transformer["wte"](input_ids) + self.transformer["wpe"](pos)
        for blk in self.transformer["h"]:
            x = blk(x)
        x = self.transformer["ln_f"](x)
        logits = F.linear(x, self.transformer["wte"].weight)          # tied LM head
        if labels is None:
            return {"logits": logits}
        # HF semantics: labels are the inputs; shift inside the model
        loss = F.cross_entropy(logits[:, :-1].reshape(-1, logits.shape[-1]).float(), labels[:, 1:].reshape(-1),
                               ignore_index=-100)
        return {"loss": loss, "logits": logits}

    def flops_per_token(self, seq_len: int) -> float:
        c = self.config
        return 6.0 * (12 * c.n_layer * c.n_embd ** 2 + c.vocab_size * c.n_embd) + 6.0 * c.n_layer * c.n_embd * seq_len
