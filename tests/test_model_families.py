"""Built-in model families against their Hugging Face originals (CPU, fp32): a checkpoint saved by `transformers` is
loaded through the executor's `get_model` and must give the same logits and the same loss the reference's
`model(**batch)` would (executors/accelerate/src/hypha/accelerate_executor/training.py:108-112)."""
import json
import os

import pytest
import torch


def _hf_checkpoint(tmp, cfgcls, cls, **extra):
    import transformers
    from safetensors.torch import save_file

    kw = dict(vocab_size=320, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
              num_key_value_heads=2, max_position_embeddings=128, rms_norm_eps=1e-5, tie_word_embeddings=False)
    if "rope_parameters" not in extra:
        kw["rope_theta"] = 10000.0
    kw.update(extra)
    cfg = getattr(transformers, cfgcls)(**kw)
    torch.manual_seed(0)
    hf = getattr(transformers, cls)(cfg).eval()
    with torch.no_grad():
        for name, p in hf.named_parameters():
            if name.endswith(".bias"):
                p.normal_(0.0, 0.5)                 # biases start at zero: make them count
            if name.endswith(("q_norm.weight", "k_norm.weight")):
                p.uniform_(0.5, 1.5)                # ... and norm weights at one
    json.dump(cfg.to_dict(), open(os.path.join(tmp, "config.json"), "w"))
    save_file({k: v.contiguous() for k, v in hf.state_dict().items()
               if not (kw.get("tie_word_embeddings") and k == "lm_head.weight")},     # a tied head is stored once
              os.path.join(tmp, "model.safetensors"))
    return hf, cfg


LLAMA3_ROPE = {"rope_parameters": {"rope_type": "llama3", "rope_theta": 10000.0, "factor": 8.0, "low_freq_factor": 1.0,
                                   "high_freq_factor": 4.0, "original_max_position_embeddings": 32}}


@pytest.mark.parametrize("family,cfgcls,cls,extra", [("llama", "LlamaConfig", "LlamaForCausalLM", {}),
                                                    ("llama", "LlamaConfig", "LlamaForCausalLM", LLAMA3_ROPE),
                                                    ("mistral", "MistralConfig", "MistralForCausalLM", {"sliding_window": 64}),
                                                    ("qwen2", "Qwen2Config", "Qwen2ForCausalLM", {}),
                                                    ("llama", "LlamaConfig", "LlamaForCausalLM", {"tie_word_embeddings": True}),
                                                    ("qwen2", "Qwen2Config", "Qwen2ForCausalLM", {"tie_word_embeddings": True}),
                                                    ("qwen3", "Qwen3Config", "Qwen3ForCausalLM", {"head_dim": 16}),
                                                    ("phi3", "Phi3Config", "Phi3ForCausalLM", {"pad_token_id": None, "sliding_window": 64})])
def test_llama_like_checkpoints_match_transformers(tmp_path, family, cfgcls, cls, extra):
    from hypha_b200.executor.model import get_model

    hf, cfg = _hf_checkpoint(str(tmp_path), cfgcls, cls, **extra)
    assert "rope_theta" not in cfg.to_dict() or True       # transformers >= 5 nests it under rope_parameters
    ours = get_model(str(tmp_path), "causal-lm", "cpu").eval()
    assert type(ours.net).__name__ == "LlamaForCausalLM" and ours.net.cfg.rope_theta == 10000.0
    assert ours.net.cfg.head_dim == 16
    assert ours.net.cfg.max_seq_len == (64 if family in ("mistral", "phi3") else 128)      # capped at the sliding window
    assert ours.net.cfg.qk_norm == (family == "qwen3") and ours.net.cfg.hf_family == family
    assert ours.net.cfg.qkv_bias == (family == "qwen2")
    ids = torch.randint(0, 320, (2, 48))
    with torch.no_grad():
        ref = hf(ids)
        got = ours(ids)["logits"].reshape(ref.logits.shape)
        assert (ref.logits - got).abs().max().item() < 1e-5
        # labels == inputs, shifted inside, the last position unscored: exactly the HF loss
        assert abs(hf(ids, labels=ids).loss.item() - ours(ids, labels=ids)["loss"].item()) < 1e-5
    # the round trip back to HF names is lossless (what a checkpoint of a finished job contains)
    sd = ours.net.to_hf_state_dict()
    tied = bool(extra.get("tie_word_embeddings"))
    assert ours.net.cfg.tie_embeddings == tied and ("lm_head.weight" in sd) == (not tied)
    for k, v in hf.state_dict().items():
        if tied and k == "lm_head.weight":
            continue                                   # stored once, as the embedding table
        assert torch.equal(sd[k].float(), v.float()), k
    # gradients: the fused q|k|v weight (and Qwen2's bias) against the three HF projections
    hf.zero_grad()
    hf(ids, labels=ids).loss.backward()
    ours.zero_grad()
    ours(ids, labels=ids)["loss"].backward()
    att = hf.model.layers[0].self_attn
    named = dict(ours.net.named_flat_parameters())
    if family == "phi3":
        want = att.qkv_proj.weight.grad
        gu = hf.model.layers[0].mlp.gate_up_proj.weight.grad
        from hypha_b200.ops import interleave_gate_up

        assert (named["layers.0.wgu"].grad - interleave_gate_up(*gu.chunk(2, 0))).abs().max().item() < 1e-5
    else:
        want = torch.cat([att.q_proj.weight.grad, att.k_proj.weight.grad, att.v_proj.weight.grad], 0)
    if family == "qwen3":
        assert att.q_norm.weight.grad.abs().max().item() > 0
        assert (named["layers.0.qnorm"].grad - att.q_norm.weight.grad).abs().max().item() < 1e-5
        assert (named["layers.0.knorm"].grad - att.k_norm.weight.grad).abs().max().item() < 1e-5
    assert (named["layers.0.wqkv"].grad - want).abs().max().item() < 1e-5
    if tied:        # one parameter, two uses: the head's gradient and the embedding's scatter-add are summed
        assert named["lm_head"] is named["embed"]
        assert (named["embed"].grad - hf.model.embed_tokens.weight.grad).abs().max().item() < 1e-5
        assert sum(p.numel() for p in ours.net.parameters()) == sum(p.numel() for p in hf.parameters())
    if family == "qwen2":
        want_b = torch.cat([att.q_proj.bias.grad, att.k_proj.bias.grad, att.v_proj.bias.grad], 0)
        assert want_b.abs().max().item() > 0 and (named["layers.0.bqkv"].grad - want_b).abs().max().item() < 1e-5


def test_unsupported_llama_variants_are_refused_loudly(tmp_path):
    from hypha_b200.executor.model import get_model

    base = {"model_type": "llama", "vocab_size": 64, "hidden_size": 32, "intermediate_size": 64, "num_hidden_layers": 1,
            "num_attention_heads": 2}
    for bad, msg in (({"rope_parameters": {"rope_type": "yarn", "rope_theta": 5e5, "factor": 8.0}}, "RoPE variant"),
                     ({"attention_bias": True}, "biases")):
        json.dump(dict(base, **bad), open(tmp_path / "config.json", "w"))
        with pytest.raises(RuntimeError, match=msg):
            get_model(str(tmp_path), "causal-lm", "cpu")


def test_gpt2_checkpoint_matches_transformers(tmp_path):
    """The plumbing config's model (BASELINE config 1: GPT-2 small): same parameter names, logits and loss as
    `transformers.GPT2LMHeadModel` (tied LM head)."""
    import transformers
    from safetensors.torch import save_file

    from hypha_b200.executor.model import get_model

    cfg = transformers.GPT2Config(vocab_size=320, n_positions=64, n_embd=64, n_layer=2, n_head=4)
    torch.manual_seed(0)
    hf = transformers.GPT2LMHeadModel(cfg).eval()
    json.dump(cfg.to_dict(), open(tmp_path / "config.json", "w"))
    save_file({k: v.contiguous() for k, v in hf.state_dict().items() if k != "lm_head.weight"},
              str(tmp_path / "model.safetensors"))
    ours = get_model(str(tmp_path), "causal-lm", "cpu").eval()
    ids = torch.randint(0, 320, (2, 32))
    with torch.no_grad():
        ref = hf(ids, labels=ids)
        got = ours(input_ids=ids, labels=ids)
    logits = got["logits"] if isinstance(got, dict) else got.logits
    loss = got["loss"] if isinstance(got, dict) else got.loss
    assert (ref.logits - logits).abs().max().item() < 1e-5 and abs(ref.loss.item() - loss.item()) < 1e-5
    assert set(ours.state_dict()) <= set(hf.state_dict())


def test_hf_label_shift():
    """Labels aligned with the inputs (the dataset convention) become per-position targets: position t <- token t+1,
    the last position unscored, dataset masks (-100) preserved; the input tensor is not modified."""
    from hypha_b200.executor.model import hf_shift_labels

    ids = torch.tensor([[5, 6, 7, 8], [1, 2, -100, 4]])
    keep = ids.clone()
    tgt, n = hf_shift_labels(ids)
    assert tgt.tolist() == [[6, 7, 8, -100], [2, -100, 4, -100]] and n == 5 and torch.equal(ids, keep)
    ref = torch.nn.functional.cross_entropy(torch.zeros(8, 10), tgt.reshape(-1), ignore_index=-100)
    assert abs(ref.item() - torch.log(torch.tensor(10.0)).item()) < 1e-6


def test_adapter_accepts_host_shifted_targets():
    """`InnerLoop.step` shifts the labels on the host (where counting the scored positions is free) and hands the
    adapter ``targets`` + ``n_valid``; that path gives the same loss as handing it Hugging Face style ``labels``."""
    from hypha_b200.executor.model import _LlamaAdapter, hf_shift_labels
    from hypha_b200.models.llama import LlamaConfig, LlamaForCausalLM

    torch.manual_seed(0)
    cfg = LlamaConfig(vocab_size=64, hidden_size=32, intermediate_size=64, num_layers=1, num_heads=2, num_kv_heads=2,
                      max_seq_len=16)
    net = _LlamaAdapter(LlamaForCausalLM(cfg, device="cpu", dtype=torch.float32))
    ids = torch.randint(0, 64, (2, 16))
    labels = ids.clone()
    labels[1, 5:9] = -100
    tgt, n = hf_shift_labels(labels)
    a = net(input_ids=ids, labels=labels)["loss"]
    b = net(input_ids=ids, targets=tgt, n_valid=n)["loss"]
    assert torch.isfinite(a) and abs(a.item() - b.item()) < 1e-6


def test_lr_schedules_match_transformers():
    """`make_lr_lambda` (folded into the fused AdamW's launch arguments) against the Hugging Face schedulers the
    reference builds (executors/accelerate/.../utils.py:90-102), step for step over the schedule's length."""
    import transformers

    from hypha_b200.parallel.diloco import make_lr_lambda

    p = torch.nn.Parameter(torch.zeros(1))

    def hf_curve(make, n):
        opt = torch.optim.SGD([p], lr=1.0)
        sch = make(opt)
        out = []
        for _ in range(n):
            out.append(opt.param_groups[0]["lr"])
            opt.step()
            sch.step()
        return out

    cases = [({"type": "cosine-with-warmup", "warmup_steps": 5, "training_steps": 50}, 51,
              lambda o: transformers.get_cosine_schedule_with_warmup(o, 5, 50)),
             ({"type": "linear-with-warmup", "warmup_steps": 5, "training_steps": 50}, 60,
              lambda o: transformers.get_linear_schedule_with_warmup(o, 5, 50)),
             ({"type": "wsd", "warmup_steps": 5, "decay_steps": 10, "training_steps": 40}, 45,
              lambda o: transformers.get_wsd_schedule(o, 5, 10, num_training_steps=40)),
             ({}, 10, lambda o: transformers.get_constant_schedule(o))]
    for cfg, n, make in cases:
        f = make_lr_lambda(cfg)
        assert max(abs(f(s) - r) for s, r in enumerate(hf_curve(make, n))) < 1e-12, cfg
    # past the end the cosine stays at zero (the Hugging Face lambda would climb again)
    f = make_lr_lambda({"type": "cosine-with-warmup", "warmup_steps": 5, "training_steps": 50})
    assert f(55) == 0.0 and f(500) == 0.0


def test_bias_gradient_follows_the_main_grad_convention():
    """The GPU form of `ops.add_bias` (an autograd Function that writes the bias gradient into the flat gradient
    buffer, overwrite or accumulate) -- its arithmetic is device independent, so it is checked here on host tensors."""
    from hypha_b200.ops.functional import _AddBias

    x = torch.randn(6, 8, requires_grad=True)
    b = torch.nn.Parameter(torch.randn(8))
    b.main_grad = torch.zeros(8)
    w = torch.randn(6, 8)
    (_AddBias.apply(x, b) * w).sum().backward()
    assert b.grad is None and torch.allclose(b.main_grad, w.sum(0)) and torch.allclose(x.grad, w)
    b.grad_accumulate = True
    (_AddBias.apply(x, b) * w).sum().backward()
    assert torch.allclose(b.main_grad, 2 * w.sum(0))
    b.grad_accumulate = False
    (_AddBias.apply(x.detach().requires_grad_(), b) * (3 * w)).sum().backward()
    assert torch.allclose(b.main_grad, 3 * w.sum(0))                  # overwritten, not added
    plain = torch.nn.Parameter(torch.randn(8))                        # no flat buffer: the ordinary `.grad`
    (_AddBias.apply(x.detach(), plain) * w).sum().backward()
    assert torch.allclose(plain.grad, w.sum(0))


def test_qwen2_preset_and_config_round_trip(tmp_path):
    from hypha_b200.models.llama import LlamaConfig, get_config

    c = get_config("qwen2.5-7b", num_layers=1, vocab_size=512)
    assert c.qkv_bias and c.head_dim * c.num_heads == c.hidden_size
    c.to_hf_json(str(tmp_path / "config.json"), seed=3)
    meta = json.load(open(tmp_path / "config.json"))
    assert meta["model_type"] == "qwen2"
    back = LlamaConfig.from_hf_dict(meta)
    assert back.qkv_bias and back.rope_theta == 1000000.0 and back.rms_eps == 1e-6
    with pytest.raises(RuntimeError, match="use_sliding_window"):
        LlamaConfig.from_hf_dict(dict(meta, use_sliding_window=True))


def test_qwen2_job_through_the_box_exports_a_transformers_checkpoint(capsys, tmp_path):
    """The whole stack on a Qwen2-shaped model (CPU): scheduler -> lease -> executor -> rounds -> export; the exported
    directory loads as `transformers.Qwen2ForCausalLM` and its q/k/v biases were trained (they start at zero)."""
    import transformers

    from hypha_b200.cli import box

    out_dir = str(tmp_path / "trained")
    rc = box.main(["run", "--model", "qwen2-tiny", "--workers", "1", "--device", "cpu", "--rounds", "2",
                   "--samples-per-round", "16", "--batch-size", "4", "--seq-len", "32", "--timeout", "240",
                   "--lr", "1e-2", "--checkpoint-dir", str(tmp_path / "ckpt"), "--export", out_dir])
    out = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert rc == 0 and out["finished"] and out["rounds"] == 2
    assert json.load(open(os.path.join(out_dir, "config.json")))["model_type"] == "qwen2"
    hf = transformers.Qwen2ForCausalLM.from_pretrained(out_dir).eval()
    with torch.no_grad():
        assert torch.isfinite(hf(torch.randint(0, 1024, (1, 16))).logits).all()
    bias = hf.model.layers[0].self_attn.q_proj.bias
    assert bias is not None and bias.abs().max().item() > 0


def test_gradients_of_plain_torch_ops_reach_the_flat_buffer():
    """`_MainGradSink`: the identity that routes a parameter's autograd gradient into ``main_grad`` (what the GPU form
    of `ops.qk_norm` wraps its norm weights in); device independent, checked on host tensors."""
    from hypha_b200.ops.functional import _MainGradSink

    w = torch.nn.Parameter(torch.randn(5))
    w.main_grad = torch.zeros(5)
    x = torch.randn(3, 5)
    (_MainGradSink.apply(w) * x).sum().backward()
    assert w.grad is None and torch.allclose(w.main_grad, x.sum(0))
    w.grad_accumulate = True
    (_MainGradSink.apply(w) * x).sum().backward()
    assert torch.allclose(w.main_grad, 2 * x.sum(0))
    plain = torch.nn.Parameter(torch.randn(5))
    (_MainGradSink.apply(plain) * x).sum().backward()
    assert torch.allclose(plain.grad, x.sum(0))


@pytest.mark.parametrize("preset,family,params_b", [("llama3-8b", "llama", 8.03), ("llama3.2-1b", "llama", 1.24),
                                                    ("mistral-7b", "mistral", 7.24), ("qwen2.5-7b", "qwen2", 7.62),
                                                    ("qwen3-8b", "qwen3", 8.19), ("phi3-mini", "phi3", 3.82)])
def test_presets_have_the_published_parameter_counts(tmp_path, preset, family, params_b):
    """The presets describe the published models: parameter count from the layout alone (no allocation), and the
    `config.json` they write is read back as the same family."""
    from hypha_b200.models.llama import LlamaConfig, get_config

    c = get_config(preset)
    h, i, v, d = c.hidden_size, c.intermediate_size, c.vocab_size, c.head_dim
    qkv = (c.num_heads + 2 * c.num_kv_heads) * d
    per_layer = qkv * h + (qkv if c.qkv_bias else 0) + (2 * d if c.qk_norm else 0) + c.num_heads * d * h + 3 * i * h + 2 * h
    total = c.num_layers * per_layer + h + v * h * (1 if c.tie_embeddings else 2)
    assert abs(total / 1e9 - params_b) < 0.01, total / 1e9
    c.to_hf_json(str(tmp_path / "config.json"))
    meta = json.load(open(tmp_path / "config.json"))
    assert meta["model_type"] == family
    back = LlamaConfig.from_hf_dict(meta)
    assert (back.qkv_bias, back.qk_norm, back.tie_embeddings, back.hf_family) == (c.qkv_bias, c.qk_norm, c.tie_embeddings, family)
    assert back.rope_scaling == c.rope_scaling and back.head_dim == c.head_dim
