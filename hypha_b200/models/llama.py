"""Llama-family causal LM built on the hypha_b200 sm_100a op library.

Design points (B200-first, not a port of HF ``LlamaForCausalLM``):

* all parameters are views into ONE flat bf16 buffer (``FlatParams``) with a
  matching flat gradient buffer, so the fused AdamW and the DiLoCo outer step are
  single launches over contiguous memory, and the buffer can be placed in
  peer-mapped symmetric memory for the NVLink outer step;
* q/k/v and gate/up projections are fused weights -> one tcgen05 GEMM each;
* residual add is fused into the following RMSNorm; RoPE runs in place on the
  fused qkv activation; the LM head is fused with the cross-entropy (chunked);
* activations can be recomputed per layer (``recompute_layers``) to trade FLOPs
  for HBM when the optimizer state of an 8B model shares one 192 GB GPU.

The reference obtains its model from ``AutoModelForCausalLM.from_pretrained``
(executors/accelerate/src/hypha/accelerate_executor/model.py:47-56); here the
architecture is local and weights are random-init or loaded from safetensors.
"""
from __future__ import annotations

import math
from dataclasses import asdict, dataclass
from typing import Dict, Iterator, List, Optional, Tuple

import torch
import torch.nn as nn

from .. import ops


@dataclass
class LlamaConfig:
    vocab_size: int = 128256
    hidden_size: int = 4096
    intermediate_size: int = 14336
    num_layers: int = 32
    num_heads: int = 32
    num_kv_heads: int = 8
    head_dim: int = 128
    rope_theta: float = 500000.0
    rope_scaling: Optional[dict] = None  # {"rope_type": "llama3", "factor", "low_freq_factor", "high_freq_factor",
                                         #  "original_max_position_embeddings"} (Llama 3.1+); None = plain RoPE
    qkv_bias: bool = False          # Qwen2: a bias on the fused q|k|v projection (and nowhere else)
    tie_embeddings: bool = False    # the LM head IS the embedding table (Llama-3.2-1B/3B, small Qwen2 models)
    qk_norm: bool = False           # Qwen3: RMSNorm over head_dim on every q and k head, before RoPE
    hf_family: str = "llama"        # whose parameter names / config.json a checkpoint uses: llama | mistral | qwen2 |
                                    # qwen3 | phi3 (phi3 stores q|k|v and gate|up fused, as this model does)
    rms_eps: float = 1e-5
    max_seq_len: int = 8192
    init_std: float = 0.02
    recompute_layers: int = 0          # first k layers recompute their activations in backward
    ce_chunk: int = 4096               # rows per fused LM-head/CE chunk
    name: str = "llama"

    @property
    def qkv_dim(self) -> int:
        return (self.num_heads + 2 * self.num_kv_heads) * self.head_dim

    def num_params(self) -> int:
        h, i, v = self.hidden_size, self.intermediate_size, self.vocab_size
        per_layer = self.qkv_dim * h + h * self.num_heads * self.head_dim + 2 * i * h + h * i + 2 * h
        return 2 * v * h + h + self.num_layers * per_layer

    def flops_per_token(self, seq_len: int) -> float:
        """Training FLOPs per token (fwd+bwd = 3x fwd; causal attention counted at half)."""
        h, i = self.hidden_size, self.intermediate_size
        dense = self.qkv_dim * h + h * self.num_heads * self.head_dim + 3 * i * h
        attn = 2 * self.num_heads * self.head_dim * seq_len / 2  # QK^T + PV, causal
        fwd = 2 * (self.num_layers * (dense + attn) + self.vocab_size * h)
        return 3.0 * fwd

    def to_dict(self) -> dict:
        return asdict(self)

    @classmethod
    def from_hf_dict(cls, meta: dict) -> "LlamaConfig":
        """The configuration of a Hugging Face ``config.json`` of a family the sm_100a kernels cover (`SM100_FAMILIES`):

        * ``llama``; ``mistral`` and ``phi3`` = the Llama block with a sliding attention window (up to the window length
          causal attention IS sliding-window attention, so the sequence length is capped there); ``phi3`` checkpoints
          store q|k|v and gate|up fused;
        * ``qwen2`` = the Llama block with a bias on q / k / v; ``qwen3`` = with an RMSNorm over every q and k head.

        RoPE settings are read from the top level (transformers < 5, hub checkpoints) or from ``rope_parameters`` /
        ``rope_scaling``; plain RoPE and the ``llama3`` frequency scaling are supported.  Anything else (other RoPE
        variants, biases on other projections, partial rotary embeddings, per-layer windows) is refused instead of
        being computed differently."""
        family = meta.get("model_type", "llama")
        if family not in SM100_FAMILIES:
            raise RuntimeError(f"model_type {family!r} is not one of the sm_100a model families {SM100_FAMILIES}")
        # qwen2 is the Llama block with a bias on q / k / v only (fixed by the architecture, not a config switch)
        if family != "qwen2" and any(meta.get(k) for k in ("attention_bias", "mlp_bias")):
            raise RuntimeError(f"{family} checkpoint with projection biases is not covered by the sm_100a model family")
        if family in ("qwen2", "qwen3") and meta.get("use_sliding_window"):
            raise RuntimeError(f"{family} with use_sliding_window (a window on the upper layers only) is not covered by "
                               "the sm_100a model family")
        if family == "phi3" and float(meta.get("partial_rotary_factor", 1.0) or 1.0) != 1.0:
            raise RuntimeError("phi3 with a partial rotary factor is not covered by the sm_100a model family")
        rope = dict(meta.get("rope_parameters") or meta.get("rope_scaling") or {})
        rope_type = rope.get("rope_type", rope.get("type", "default")) or "default"
        if rope_type not in ("default", "llama3"):
            raise RuntimeError(f"RoPE variant {rope_type!r} is not covered by the sm_100a model family "
                               f"(plain RoPE and llama3 scaling only)")
        theta = meta.get("rope_theta", rope.get("rope_theta"))
        if theta is None:
            theta = 500000.0 if family == "llama" else 10000.0
        scaling = None
        if rope_type == "llama3":
            scaling = {"rope_type": "llama3", "factor": float(rope["factor"]),
                       "low_freq_factor": float(rope.get("low_freq_factor", 1.0)),
                       "high_freq_factor": float(rope.get("high_freq_factor", 4.0)),
                       "original_max_position_embeddings": int(rope.get("original_max_position_embeddings", 8192))}
        max_pos = int(meta.get("max_position_embeddings", 8192))
        if family in ("mistral", "phi3") and meta.get("sliding_window"):
            max_pos = min(max_pos, int(meta["sliding_window"]))
        heads = int(meta["num_attention_heads"])
        return cls(vocab_size=int(meta["vocab_size"]), hidden_size=int(meta["hidden_size"]),
                   intermediate_size=int(meta["intermediate_size"]), num_layers=int(meta["num_hidden_layers"]),
                   num_heads=heads, num_kv_heads=int(meta.get("num_key_value_heads") or heads),
                   head_dim=int(meta.get("head_dim") or int(meta["hidden_size"]) // heads), rope_theta=float(theta),
                   rope_scaling=scaling, qkv_bias=family == "qwen2", qk_norm=family == "qwen3", hf_family=family,
                   tie_embeddings=bool(meta.get("tie_word_embeddings", False)),
                   rms_eps=float(meta.get("rms_norm_eps", 1e-6 if family in ("qwen2", "qwen3") else 1e-5)),
                   max_seq_len=max_pos,
                   name=meta.get("name", family))

    def to_hf_json(self, path: str, seed: int = 0) -> None:
        """Write a Hugging-Face style ``config.json`` (what ``executor/model.py`` and the collective
        executor read back); ``seed`` fixes the random init every worker of a job starts from."""
        import json

        family = "qwen2" if self.qkv_bias else "qwen3" if self.qk_norm else self.hf_family
        arch = {"llama": "LlamaForCausalLM", "mistral": "MistralForCausalLM", "qwen2": "Qwen2ForCausalLM",
                "qwen3": "Qwen3ForCausalLM", "phi3": "Phi3ForCausalLM"}[family]
        json.dump({"model_type": family, "architectures": [arch], "name": self.name,
                   "vocab_size": self.vocab_size, "hidden_size": self.hidden_size,
                   "intermediate_size": self.intermediate_size, "num_hidden_layers": self.num_layers,
                   "num_attention_heads": self.num_heads, "num_key_value_heads": self.num_kv_heads,
                   "head_dim": self.head_dim, "rope_theta": self.rope_theta, "rms_norm_eps": self.rms_eps,
                   "max_position_embeddings": self.max_seq_len, "seed": seed,
                   "tie_word_embeddings": self.tie_embeddings,
                   **({"rope_scaling": dict(self.rope_scaling)} if self.rope_scaling else {})}, open(path, "w"))


# config.json `model_type`s that run on the sm_100a kernels through `LlamaForCausalLM` (`LlamaConfig.from_hf_dict`)
SM100_FAMILIES = ("llama", "mistral", "qwen2", "qwen3", "phi3")

PRESETS: Dict[str, LlamaConfig] = {
    "llama3-8b": LlamaConfig(name="llama3-8b"),
    "llama3-1b": LlamaConfig(vocab_size=128256, hidden_size=2048, intermediate_size=8192,
                             num_layers=16, num_heads=32, num_kv_heads=8, head_dim=64,
                             name="llama3-1b"),
    "llama-tiny": LlamaConfig(vocab_size=1024, hidden_size=256, intermediate_size=512,
                              num_layers=2, num_heads=4, num_kv_heads=2, head_dim=64,
                              max_seq_len=256, ce_chunk=128, name="llama-tiny"),
    "qwen2.5-7b": LlamaConfig(vocab_size=152064, hidden_size=3584, intermediate_size=18944, num_layers=28,
                              num_heads=28, num_kv_heads=4, head_dim=128, rope_theta=1000000.0, qkv_bias=True,
                              rms_eps=1e-6, max_seq_len=32768, hf_family="qwen2", name="qwen2.5-7b"),
    "llama3.2-1b": LlamaConfig(vocab_size=128256, hidden_size=2048, intermediate_size=8192, num_layers=16,
                               num_heads=32, num_kv_heads=8, head_dim=64, tie_embeddings=True, max_seq_len=131072,
                               rope_scaling={"rope_type": "llama3", "factor": 32.0, "low_freq_factor": 1.0,
                                             "high_freq_factor": 4.0, "original_max_position_embeddings": 8192},
                               name="llama3.2-1b"),
    "mistral-7b": LlamaConfig(vocab_size=32000, hidden_size=4096, intermediate_size=14336, num_layers=32,
                              num_heads=32, num_kv_heads=8, head_dim=128, rope_theta=10000.0, max_seq_len=4096,
                              hf_family="mistral", name="mistral-7b"),          # v0.1: 4096-token sliding window
    "qwen3-8b": LlamaConfig(vocab_size=151936, hidden_size=4096, intermediate_size=12288, num_layers=36,
                            num_heads=32, num_kv_heads=8, head_dim=128, rope_theta=1000000.0, qk_norm=True,
                            rms_eps=1e-6, max_seq_len=40960, hf_family="qwen3", name="qwen3-8b"),
    "phi3-mini": LlamaConfig(vocab_size=32064, hidden_size=3072, intermediate_size=8192, num_layers=32,
                             num_heads=32, num_kv_heads=32, head_dim=96, rope_theta=10000.0, max_seq_len=2047,
                             hf_family="phi3", name="phi3-mini"),               # 4k variant: 2047-token window
    "qwen2-tiny": LlamaConfig(vocab_size=1024, hidden_size=256, intermediate_size=512, num_layers=2, num_heads=4,
                              num_kv_heads=2, head_dim=64, rope_theta=1000000.0, qkv_bias=True, rms_eps=1e-6,
                              max_seq_len=256, ce_chunk=128, hf_family="qwen2", name="qwen2-tiny"),
    "llama-small": LlamaConfig(vocab_size=32000, hidden_size=1024, intermediate_size=4096,
                               num_layers=4, num_heads=8, num_kv_heads=4, head_dim=128,
                               max_seq_len=2048, ce_chunk=2048, name="llama-small"),
}


def get_config(name: str, **overrides) -> LlamaConfig:
    cfg = PRESETS[name]
    d = cfg.to_dict()
    d.update(overrides)
    return LlamaConfig(**d)


class FlatParams:
    """One flat bf16 parameter buffer + one flat bf16 gradient buffer.

    ``alloc`` may be supplied to place the parameter buffer in peer-mapped
    symmetric memory (see ``hypha_b200.parallel.peer_memory``)."""

    ALIGN = 64  # elements; keeps every tensor 128-byte aligned for TMA

    def __init__(self, shapes: List[Tuple[str, Tuple[int, ...]]], device, dtype=torch.bfloat16,
                 alloc=None, grad_alloc=None, with_grad: bool = True):
        self.layout: Dict[str, Tuple[int, Tuple[int, ...]]] = {}
        off = 0
        for name, shape in shapes:
            n = math.prod(shape)
            self.layout[name] = (off, shape)
            off += (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.numel = (off + 1023) // 1024 * 1024
        if alloc is not None:
            self.data = alloc(self.numel, dtype)
        else:
            self.data = torch.zeros(self.numel, dtype=dtype, device=device)
        if not with_grad:
            self.grad = None
        elif grad_alloc is not None:
            self.grad = grad_alloc(self.numel, dtype)
        else:
            self.grad = torch.zeros(self.numel, dtype=dtype, device=device)

    def view(self, name: str, of: Optional[torch.Tensor] = None) -> torch.Tensor:
        off, shape = self.layout[name]
        base = self.data if of is None else of
        return base[off: off + math.prod(shape)].view(shape)

    def named_views(self) -> Iterator[Tuple[str, torch.Tensor]]:
        for name in self.layout:
            yield name, self.view(name)


class _CheckpointPair(torch.autograd.Function):
    """Recompute-in-backward for a ``(delta, residual) -> (delta, residual)`` layer
    (``residual`` may be ``None`` for the first layer)."""

    @staticmethod
    def forward(ctx, fn, delta, residual):
        ctx.fn = fn
        ctx.has_res = residual is not None
        if ctx.has_res:
            ctx.save_for_backward(delta, residual)
        else:
            ctx.save_for_backward(delta)
        with torch.no_grad():
            d, r = fn(delta, residual)
        if r is delta:          # first layer: the residual stream IS the input; never alias outputs
            r = r.clone()
        return d, r

    @staticmethod
    def backward(ctx, dd, dr):
        saved = ctx.saved_tensors
        delta = saved[0].detach().requires_grad_(True)
        residual = saved[1].detach().requires_grad_(True) if ctx.has_res else None
        with torch.enable_grad():
            d, r = ctx.fn(delta, residual)
            torch.autograd.backward([d, r], [dd, dr])
        return None, delta.grad, (residual.grad if ctx.has_res else None)


class LlamaBlock(nn.Module):
    def __init__(self, cfg: LlamaConfig, wqkv, wo, wgu, wdown, norm1, norm2, bqkv=None, qnorm=None, knorm=None):
        super().__init__()
        self.cfg = cfg
        self.wqkv, self.wo, self.wgu, self.wdown = wqkv, wo, wgu, wdown
        self.norm1, self.norm2 = norm1, norm2
        self.bqkv = bqkv
        self.qnorm, self.knorm = qnorm, knorm

    def forward(self, delta, residual, cos, sin, batch: int, seq: int):
        c = self.cfg
        qkv, res = ops.add_norm_linear(delta, residual, self.norm1, self.wqkv, c.rms_eps)
        if self.bqkv is not None:
            qkv = ops.add_bias(qkv, self.bqkv)
        if self.qnorm is not None:
            qkv = ops.qk_norm(qkv, self.qnorm, self.knorm, c.num_heads, c.num_kv_heads, c.head_dim, c.rms_eps)
        qkv = ops.rope_qkv(qkv, cos, sin, seq, c.num_heads + c.num_kv_heads, c.head_dim, grad_inplace=True)
        att = ops.attention_qkv(qkv, batch, seq, c.num_heads, c.num_kv_heads, c.head_dim)
        o = ops.linear(att, self.wo)
        # the gate_up GEMM's epilogue emits silu(gate) * up next to gate|up (interleaved columns)
        act = torch.empty((o.shape[0], c.intermediate_size), dtype=o.dtype, device=o.device) if o.is_cuda else None
        gu, res = ops.add_norm_linear(o, res, self.norm2, self.wgu, c.rms_eps, swiglu_out=act)
        down = ops.swiglu_linear(gu, self.wdown, act=act)
        return down, res


class LlamaForCausalLM(nn.Module):
    """Causal LM; ``forward(input_ids, labels)`` returns the mean token loss."""

    def __init__(self, cfg: LlamaConfig, device="cuda", dtype=torch.bfloat16, alloc=None,
                 grad_alloc=None, seed: int = 0, init: bool = True):
        super().__init__()
        self.cfg = cfg
        h, i, v = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size
        shapes: List[Tuple[str, Tuple[int, ...]]] = [("embed", (v, h))]
        for l in range(cfg.num_layers):
            shapes += [
                (f"layers.{l}.norm1", (h,)),
                (f"layers.{l}.wqkv", (cfg.qkv_dim, h)),
                *([(f"layers.{l}.bqkv", (cfg.qkv_dim,))] if cfg.qkv_bias else []),
                *([(f"layers.{l}.qnorm", (cfg.head_dim,)), (f"layers.{l}.knorm", (cfg.head_dim,))] if cfg.qk_norm else []),
                (f"layers.{l}.wo", (h, cfg.num_heads * cfg.head_dim)),
                (f"layers.{l}.norm2", (h,)),
                (f"layers.{l}.wgu", (2 * i, h)),
                (f"layers.{l}.wdown", (h, i)),
            ]
        shapes += [("final_norm", (h,))] + ([] if cfg.tie_embeddings else [("lm_head", (v, h))])
        self.flat = FlatParams(shapes, device, dtype, alloc=alloc, grad_alloc=grad_alloc)
        self._params: Dict[str, nn.Parameter] = {}
        for name, view in self.flat.named_views():
            p = nn.Parameter(view, requires_grad=True)
            p.main_grad = self.flat.view(name, of=self.flat.grad)
            p.grad_accumulate = False
            self._params[name] = p
            self.register_parameter(name.replace(".", "_"), p)
        P = self._params
        if cfg.tie_embeddings:
            # ONE parameter used twice: the head's weight gradient (written by the fused LM-head + cross-entropy op
            # in overwrite mode) and the embedding's scatter-add meet in the same slice of the flat gradient buffer,
            # so the embedding backward must add to what it finds (`ops.functional._Embedding.backward`)
            P["lm_head"] = P["embed"]
            P["embed"].tied_head = True
        self.blocks = nn.ModuleList([
            LlamaBlock(cfg, P[f"layers.{l}.wqkv"], P[f"layers.{l}.wo"], P[f"layers.{l}.wgu"],
                       P[f"layers.{l}.wdown"], P[f"layers.{l}.norm1"], P[f"layers.{l}.norm2"],
                       P.get(f"layers.{l}.bqkv"), P.get(f"layers.{l}.qnorm"), P.get(f"layers.{l}.knorm"))
            for l in range(cfg.num_layers)
        ])
        self._rope_cache: Dict[Tuple[int, str], Tuple[torch.Tensor, torch.Tensor]] = {}
        # set by the trainer: callable(begin, end) invoked from backward as soon as the gradients of
        # the flat parameter range [begin, end) are complete (optimizer-in-backward overlap)
        self.grad_ready_hook = None
        L = self.flat.layout
        starts = [L[f"layers.{l}.norm1"][0] for l in range(cfg.num_layers)] + [L["final_norm"][0]]
        self.layer_ranges = [(starts[l], starts[l + 1]) for l in range(cfg.num_layers)]
        self.head_range = (L["final_norm"][0], self.flat.numel)
        self.embed_range = (0, starts[0])
        if init:
            self.reset_parameters(seed)

    # -- parameters ---------------------------------------------------------------------
    @torch.no_grad()
    def reset_parameters(self, seed: int = 0) -> None:
        g = torch.Generator(device=self.flat.data.device)
        g.manual_seed(seed)
        cfg = self.cfg
        for name, view in self.flat.named_views():
            if "norm" in name:
                view.fill_(1.0)
            elif name.endswith("bqkv"):
                view.zero_()
            else:
                std = cfg.init_std
                if name.endswith("wo") or name.endswith("wdown"):
                    std = cfg.init_std / math.sqrt(2 * cfg.num_layers)
                view.normal_(0.0, std, generator=g)

    def named_flat_parameters(self) -> Dict[str, nn.Parameter]:
        return dict(self._params)

    def set_grad_accumulate(self, flag: bool) -> None:
        for p in self._params.values():
            p.grad_accumulate = flag

    def _rope(self, seq: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
        key = (seq, str(device))
        if key not in self._rope_cache:
            self._rope_cache[key] = ops.rope_tables(seq, self.cfg.head_dim, self.cfg.rope_theta, device,
                                                    scaling=self.cfg.rope_scaling)
        return self._rope_cache[key]

    # -- forward ------------------------------------------------------------------------
    def forward(self, input_ids: torch.Tensor, labels: Optional[torch.Tensor] = None,
                ignore_index: Optional[int] = None, n_valid: Optional[int] = None):
        """Logits ``[B, S, V]``, or with ``labels`` (already aligned with the positions: ``labels[b, t]`` is the
        target of position ``t``) the mean cross-entropy.  ``ignore_index`` masks targets; ``n_valid`` = number of
        unmasked targets when the caller knows it (no host synchronisation)."""
        cfg = self.cfg
        batch, seq = input_ids.shape
        cos, sin = self._rope(seq, input_ids.device)
        P = self._params
        delta = ops.embedding(input_ids, P["embed"])
        residual = None
        hook = self.grad_ready_hook if torch.is_grad_enabled() else None
        for l, blk in enumerate(self.blocks):
            if hook is not None and delta.requires_grad:
                rng = self.layer_ranges[l]
                delta.register_hook(lambda g, rng=rng: hook(*rng))      # fires when layer l's backward is done
            if l < cfg.recompute_layers and torch.is_grad_enabled():
                fn = lambda d, r, blk=blk: blk(d, r, cos, sin, batch, seq)
                delta, residual = _CheckpointPair.apply(fn, delta, residual)
            else:
                delta, residual = blk(delta, residual, cos, sin, batch, seq)
        if hook is not None and delta.requires_grad:
            delta.register_hook(lambda g: hook(*self.head_range))       # final norm + LM head grads complete
        hN, _ = ops.add_rmsnorm(delta, residual, P["final_norm"], cfg.rms_eps)
        if labels is None:
            return ops.linear(hN, P["lm_head"]).view(batch, seq, -1)
        return ops.linear_cross_entropy(hN, P["lm_head"], labels.reshape(-1), chunk=cfg.ce_chunk,
                                        ignore_index=ignore_index, n_valid=n_valid)

    # -- HF interop -----------------------------------------------------------------------
    @torch.no_grad()
    def load_hf_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        """Load a HuggingFace ``LlamaForCausalLM`` state dict (fuses q/k/v and gate/up)."""
        P = self._params
        P["embed"].copy_(sd["model.embed_tokens.weight"])
        P["final_norm"].copy_(sd["model.norm.weight"])
        if not self.cfg.tie_embeddings:
            P["lm_head"].copy_(sd.get("lm_head.weight", sd["model.embed_tokens.weight"]))
        for l in range(self.cfg.num_layers):
            pre = f"model.layers.{l}."
            if self.cfg.hf_family == "phi3":                 # already fused: [q; k; v] and [gate; up] row blocks
                P[f"layers.{l}.wqkv"].copy_(sd[pre + "self_attn.qkv_proj.weight"])
                gate, up = sd[pre + "mlp.gate_up_proj.weight"].chunk(2, 0)
            else:
                P[f"layers.{l}.wqkv"].copy_(torch.cat([sd[pre + "self_attn.q_proj.weight"],
                                                       sd[pre + "self_attn.k_proj.weight"],
                                                       sd[pre + "self_attn.v_proj.weight"]], 0))
                gate, up = sd[pre + "mlp.gate_proj.weight"], sd[pre + "mlp.up_proj.weight"]
            if self.cfg.qkv_bias:
                P[f"layers.{l}.bqkv"].copy_(torch.cat([sd[pre + "self_attn.q_proj.bias"],
                                                       sd[pre + "self_attn.k_proj.bias"],
                                                       sd[pre + "self_attn.v_proj.bias"]], 0))
            if self.cfg.qk_norm:
                P[f"layers.{l}.qnorm"].copy_(sd[pre + "self_attn.q_norm.weight"])
                P[f"layers.{l}.knorm"].copy_(sd[pre + "self_attn.k_norm.weight"])
            P[f"layers.{l}.wo"].copy_(sd[pre + "self_attn.o_proj.weight"])
            P[f"layers.{l}.wgu"].copy_(ops.interleave_gate_up(gate, up))
            P[f"layers.{l}.wdown"].copy_(sd[pre + "mlp.down_proj.weight"])
            P[f"layers.{l}.norm1"].copy_(sd[pre + "input_layernorm.weight"])
            P[f"layers.{l}.norm2"].copy_(sd[pre + "post_attention_layernorm.weight"])

    def to_hf_state_dict(self) -> Dict[str, torch.Tensor]:
        c, P = self.cfg, self._params
        qd, kd = c.num_heads * c.head_dim, c.num_kv_heads * c.head_dim
        sd = {"model.embed_tokens.weight": P["embed"].detach(), "model.norm.weight": P["final_norm"].detach()}
        if not c.tie_embeddings:                    # a tied head is not stored: `transformers` re-ties it on load
            sd["lm_head.weight"] = P["lm_head"].detach()
        for l in range(c.num_layers):
            pre = f"model.layers.{l}."
            wqkv, wgu = P[f"layers.{l}.wqkv"].detach(), P[f"layers.{l}.wgu"].detach()
            gate, up = ops.split_gate_up(wgu)
            if c.hf_family == "phi3":
                sd[pre + "self_attn.qkv_proj.weight"] = wqkv
                sd[pre + "mlp.gate_up_proj.weight"] = torch.cat([gate, up], 0)
            else:
                sd[pre + "self_attn.q_proj.weight"] = wqkv[:qd]
                sd[pre + "self_attn.k_proj.weight"] = wqkv[qd: qd + kd]
                sd[pre + "self_attn.v_proj.weight"] = wqkv[qd + kd:]
                sd[pre + "mlp.gate_proj.weight"], sd[pre + "mlp.up_proj.weight"] = gate, up
            if c.qkv_bias:
                b = P[f"layers.{l}.bqkv"].detach()
                sd[pre + "self_attn.q_proj.bias"], sd[pre + "self_attn.k_proj.bias"] = b[:qd], b[qd: qd + kd]
                sd[pre + "self_attn.v_proj.bias"] = b[qd + kd:]
            if c.qk_norm:
                sd[pre + "self_attn.q_norm.weight"] = P[f"layers.{l}.qnorm"].detach()
                sd[pre + "self_attn.k_norm.weight"] = P[f"layers.{l}.knorm"].detach()
            sd[pre + "self_attn.o_proj.weight"] = P[f"layers.{l}.wo"].detach()
            sd[pre + "mlp.down_proj.weight"] = P[f"layers.{l}.wdown"].detach()
            sd[pre + "input_layernorm.weight"] = P[f"layers.{l}.norm1"].detach()
            sd[pre + "post_attention_layernorm.weight"] = P[f"layers.{l}.norm2"].detach()
        return sd
