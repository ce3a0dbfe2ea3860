"""`hypha_b200.export` / `hypha-box export`: a job checkpoint of any of the three kinds becomes a model directory that
`transformers` loads, holding exactly the global model the job had reached."""
import json
import os

import pytest
import torch


def _gpt2_dir(tmp):
    import transformers

    cfg = transformers.GPT2Config(vocab_size=320, n_positions=64, n_embd=64, n_layer=2, n_head=4)
    d = dict(cfg.to_dict(), seed=7)
    os.makedirs(tmp, exist_ok=True)
    json.dump(d, open(os.path.join(tmp, "config.json"), "w"))
    json.dump({"dummy": True}, open(os.path.join(tmp, "tokenizer_config.json"), "w"))
    return cfg


def test_parameter_server_checkpoint_exports_theta0_plus_update(tmp_path):
    import transformers
    from safetensors.torch import save_file

    from hypha_b200.executor.model import get_model
    from hypha_b200.export import ExportError, export_checkpoint
    from hypha_b200.worker.jobs import save_job_checkpoint

    mdir, ck, out = str(tmp_path / "model"), str(tmp_path / "ckpt"), str(tmp_path / "out")
    _gpt2_dir(mdir)
    theta0 = {k: v.clone() for k, v in get_model(mdir, "causal-lm", "cpu").state_dict().items()}
    torch.manual_seed(1)
    U = {k: 0.01 * torch.randn_like(v) for k, v in theta0.items() if v.dtype.is_floating_point and k != "lm_head.weight"}
    save_file(U, str(tmp_path / "cum"))
    save_file({k: torch.zeros_like(v) for k, v in U.items()}, str(tmp_path / "mom"))
    os.makedirs(ck)
    with pytest.raises(ExportError, match="no complete checkpoint"):
        export_checkpoint(ck, mdir, out)
    save_job_checkpoint(ck, 3, str(tmp_path / "cum"), str(tmp_path / "mom"))
    info = export_checkpoint(ck, mdir, out)
    assert info["round"] == 3 and info["kind"] == "parameter-server" and info["checkpoint_tensors_used"] == len(U)
    assert os.path.isfile(os.path.join(out, "config.json")) and os.path.isfile(os.path.join(out, "tokenizer_config.json"))
    hf = transformers.GPT2LMHeadModel.from_pretrained(out).eval()
    sd = hf.state_dict()
    for k, u in U.items():
        assert torch.allclose(sd[k], theta0[k] + u, atol=1e-6), k
    assert sd["lm_head.weight"].data_ptr() == sd["transformer.wte.weight"].data_ptr()      # re-tied on load
    with torch.no_grad():
        assert torch.isfinite(hf(torch.randint(0, 320, (1, 16))).logits).all()
    with pytest.raises(ExportError, match="no checkpoint of round 9"):
        export_checkpoint(ck, mdir, out, round_no=9)
    with pytest.raises(ExportError, match="dtype"):
        export_checkpoint(ck, mdir, out, dtype="int4")


def test_collective_state_exports_theta_prev(tmp_path):
    import transformers
    from safetensors.torch import save_file

    from hypha_b200.executor.model import get_model
    from hypha_b200.export import export_checkpoint
    from hypha_b200.worker.jobs import publish_checkpoint

    mdir, ck, out = str(tmp_path / "model"), str(tmp_path / "ckpt"), str(tmp_path / "out")
    _gpt2_dir(mdir)
    sd = get_model(mdir, "causal-lm", "cpu").state_dict()
    torch.manual_seed(2)
    prev = {k: torch.randn_like(v) * 0.02 for k, v in sd.items() if v.dtype.is_floating_point and k != "lm_head.weight"}
    state = {f"prev/{k}": v for k, v in prev.items()}
    state.update({f"m/{k}": torch.zeros_like(v) for k, v in prev.items()})
    save_file(state, str(tmp_path / "cs"))
    publish_checkpoint(ck, 5, {"collective_state": str(tmp_path / "cs")})
    info = export_checkpoint(ck, mdir, out, dtype="bf16")
    assert info["kind"] == "collective" and info["round"] == 5
    hf = transformers.GPT2LMHeadModel.from_pretrained(out)
    for k, v in prev.items():
        assert torch.allclose(hf.state_dict()[k].float(), v.bfloat16().float()), k


@pytest.mark.parametrize("with_shards", [True, False])
def test_fused_engine_checkpoint_exports_the_llama_weights(tmp_path, with_shards):
    """Flat bf16 parameters + one fp32 theta_outer shard per rank (what `_TrainerEngine.save_checkpoint` writes): the
    export takes the fp32 shards when they are complete, the bf16 vector otherwise, and the result loads into
    `transformers.LlamaForCausalLM` with the same logits as our model on that vector."""
    import transformers
    from safetensors.torch import save_file

    from hypha_b200.export import export_checkpoint
    from hypha_b200.models.llama import LlamaConfig, LlamaForCausalLM
    from hypha_b200.worker.jobs import publish_sharded_checkpoint

    mdir, ck, out = str(tmp_path / "model"), str(tmp_path / "ckpt"), str(tmp_path / "out")
    os.makedirs(mdir)
    cfg = LlamaConfig(vocab_size=320, hidden_size=64, intermediate_size=128, num_layers=2, num_heads=4, num_kv_heads=2,
                      max_seq_len=64)
    cfg.to_hf_json(os.path.join(mdir, "config.json"), seed=3)
    ours = LlamaForCausalLM(cfg, device="cpu", dtype=torch.float32, seed=11)
    flat32 = ours.flat.data.clone()
    n, world = flat32.numel(), 2
    os.makedirs(ck)
    half = (n // 2 // 8) * 8
    bounds = [(0, half), (half, n - half)]
    names = ["collective_params"] + [f"outer_shard_{r}" for r in range(world)]
    if with_shards:
        for r, (b, ln) in enumerate(bounds):
            save_file({"theta_outer": flat32[b:b + ln].clone(), "momentum": torch.zeros(ln),
                       "meta": torch.tensor([4, b, ln, world])}, str(tmp_path / f"s{r}"))
            publish_sharded_checkpoint(ck, 4, f"outer_shard_{r}", str(tmp_path / f"s{r}"))
    else:
        names = ["collective_params"]
    save_file({"params": flat32.bfloat16()}, str(tmp_path / "p"))
    publish_sharded_checkpoint(ck, 4, "collective_params", str(tmp_path / "p"), expect=names)
    info = export_checkpoint(ck, mdir, out)
    assert info["kind"] == "collective-fused" and info["round"] == 4
    hf = transformers.LlamaForCausalLM.from_pretrained(out).eval()
    if not with_shards:
        ours.flat.data.copy_(flat32.bfloat16().float())
    ids = torch.randint(0, 320, (2, 24))
    with torch.no_grad():
        assert (hf(ids).logits - ours(ids).reshape(2, 24, -1)).abs().max().item() < 1e-4
    exact = sum(torch.equal(v, ours.to_hf_state_dict()[k]) for k, v in hf.state_dict().items())
    assert exact == len(hf.state_dict())


def test_parameter_server_update_in_flat_layout_names(tmp_path):
    """GPU executors of a parameter-server job report their pseudo-gradients under the names of the flat sm_100a
    layout (`layers.0.wqkv`, ...); the export applies such an update to the Llama model as well."""
    import transformers
    from safetensors.torch import save_file

    from hypha_b200.executor.model import get_model
    from hypha_b200.export import ExportError, export_checkpoint
    from hypha_b200.models.llama import LlamaConfig
    from hypha_b200.worker.jobs import save_job_checkpoint

    mdir, ck, out = str(tmp_path / "model"), str(tmp_path / "ckpt"), str(tmp_path / "out")
    os.makedirs(mdir)
    os.makedirs(ck)
    cfg = LlamaConfig(vocab_size=320, hidden_size=64, intermediate_size=128, num_layers=1, num_heads=4, num_kv_heads=2,
                      max_seq_len=64)
    cfg.to_hf_json(os.path.join(mdir, "config.json"), seed=3)
    base = get_model(mdir, "causal-lm", "cpu")
    before = {k: v.clone() for k, v in base.net.to_hf_state_dict().items()}
    U = {name: torch.full(shape, 0.5) for name, (_, shape) in base.net.flat.layout.items()}
    save_file(U, str(tmp_path / "cum"))
    save_file({k: torch.zeros_like(v) for k, v in U.items()}, str(tmp_path / "mom"))
    save_job_checkpoint(ck, 1, str(tmp_path / "cum"), str(tmp_path / "mom"))
    info = export_checkpoint(ck, mdir, out)
    assert info["checkpoint_tensors_used"] == len(U)
    hf = transformers.LlamaForCausalLM.from_pretrained(out)
    for k, v in hf.state_dict().items():
        assert torch.allclose(v, before[k].float() + 0.5, atol=1e-6), k
    save_file({"nonsense": torch.zeros(3)}, str(tmp_path / "cum"))
    save_job_checkpoint(ck, 2, str(tmp_path / "cum"), str(tmp_path / "mom"))
    with pytest.raises(ExportError, match="match neither"):
        export_checkpoint(ck, mdir, out)


def test_box_fine_tunes_a_sharded_hf_checkpoint_and_exports_it(tmp_path, capsys):
    """`hypha-box run --model-dir`: the job starts from the weights of a Hugging Face directory -- here published as
    two shards with an index, the way large checkpoints are -- and `--export` writes the result back: after one
    round at a tiny learning rate the exported weights are the originals plus a small step, not a re-initialisation."""
    import transformers
    from safetensors.torch import save_file

    from hypha_b200.cli import box
    from hypha_b200.executor.model import get_model, weight_files

    mdir = tmp_path / "tiny-llama"
    mdir.mkdir()
    cfg = transformers.LlamaConfig(vocab_size=320, hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                                   num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=64,
                                   rms_norm_eps=1e-5, tie_word_embeddings=False)
    torch.manual_seed(5)
    hf = transformers.LlamaForCausalLM(cfg)
    json.dump(cfg.to_dict(), open(mdir / "config.json", "w"))
    sd = {k: v.contiguous() for k, v in hf.state_dict().items()}
    keys = sorted(sd)
    parts = {"model-00001-of-00002.safetensors": keys[: len(keys) // 2], "model-00002-of-00002.safetensors": keys[len(keys) // 2:]}
    for name, ks in parts.items():
        save_file({k: sd[k] for k in ks}, str(mdir / name))
    json.dump({"metadata": {}, "weight_map": {k: n for n, ks in parts.items() for k in ks}},
              open(mdir / "model.safetensors.index.json", "w"))
    json.dump({"model_max_length": 64}, open(mdir / "tokenizer_config.json", "w"))
    assert [os.path.basename(f) for f in weight_files(str(mdir))] == sorted(parts)
    assert box.model_dir_files(str(mdir)) == ["config.json", "model.safetensors.index.json", "tokenizer_config.json"] + sorted(parts)
    loaded = get_model(str(mdir), "causal-lm", "cpu")
    assert torch.equal(loaded.net.to_hf_state_dict()["model.embed_tokens.weight"], sd["model.embed_tokens.weight"])
    out = tmp_path / "tuned"
    rc = box.main(["run", "--model-dir", str(mdir), "--workers", "1", "--device", "cpu", "--rounds", "1",
                   "--samples-per-round", "8", "--batch-size", "4", "--seq-len", "32", "--lr", "1e-5", "--timeout", "240",
                   "--checkpoint-dir", str(tmp_path / "ck"), "--export", str(out)])
    res = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert rc == 0 and res["finished"] and res["exported"]["round"] == 1
    tuned = transformers.LlamaForCausalLM.from_pretrained(str(out)).state_dict()
    worst = max((tuned[k].float() - v.float()).abs().max().item() for k, v in sd.items())
    assert 0 < worst < 5e-3                       # the original weights, moved a little
    assert os.path.isfile(out / "tokenizer_config.json")
    # an index that names a shard which is not there is an error, not a silent random initialisation
    os.unlink(mdir / "model-00002-of-00002.safetensors")
    with pytest.raises(FileNotFoundError, match="not in"):
        get_model(str(mdir), "causal-lm", "cpu")


def test_eval_reports_the_hf_loss_over_a_slice_directory(tmp_path, capsys):
    """`hypha-box eval`: token-weighted mean of the Hugging Face loss over the slices, equal to what `transformers`
    computes for the same checkpoint and data."""
    import math

    import transformers
    from safetensors.torch import save_file

    from hypha_b200.cli import box
    from hypha_b200.executor.dataset import write_synthetic_token_slices

    mdir = tmp_path / "m"
    mdir.mkdir()
    cfg = transformers.LlamaConfig(vocab_size=320, hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                                   num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=64,
                                   rms_norm_eps=1e-5, tie_word_embeddings=False)
    torch.manual_seed(9)
    hf = transformers.LlamaForCausalLM(cfg).eval()
    json.dump(cfg.to_dict(), open(mdir / "config.json", "w"))
    save_file({k: v.contiguous() for k, v in hf.state_dict().items()}, str(mdir / "model.safetensors"))
    files = write_synthetic_token_slices(str(tmp_path / "ds"), num_slices=2, samples_per_slice=5, seq_len=32, vocab_size=320)
    assert box.main(["eval", "--model", str(mdir), "--dataset-path", str(tmp_path / "ds"), "--device", "cpu",
                     "--batch-size", "4"]) == 0
    got = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    from safetensors.torch import load_file

    ids = torch.cat([load_file(f)["input_ids"] for f in files])
    with torch.no_grad():
        ref = hf(ids, labels=ids).loss.item()
    assert got["samples"] == 10 and got["tokens_scored"] == 10 * 31
    assert abs(got["loss"] - ref) < 1e-4 and abs(got["perplexity"] - math.exp(ref)) < 1e-1
    assert box.main(["eval", "--model", str(mdir), "--dataset-path", str(tmp_path / "ds"), "--device", "cpu",
                     "--max-samples", "3"]) == 0
    assert json.loads(capsys.readouterr().out.strip().splitlines()[-1])["samples"] == 3
    (tmp_path / "empty").mkdir()
    assert box.main(["eval", "--model", str(mdir), "--dataset-path", str(tmp_path / "empty"), "--device", "cpu"]) == 2
