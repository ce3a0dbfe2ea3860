"""The reference arm of ``bench.py`` (``--impl reference`` / ``--impl stock``).

What runs here, and why it is a re-creation rather than the unmodified reference
------------------------------------------------------------------------------
The reference (hypha-space/hypha @ 76fe4718) is a Rust/libp2p control plane plus one Python
executor package, ``executors/accelerate``.  The Rust half cannot be built offline (no cargo/rustc,
libp2p patched from a git fork).  The Python half *does* install into ``baseline/_ref`` (from a /tmp
copy whose ``[build-system]`` is switched from the absent ``uv_build`` to setuptools; the sources are
byte-identical), but its entry point cannot be imported in this image:

* ``training.py`` needs ``accelerate==1.10.1`` -- not installed, not in /opt/wheelhouse;
* ``model.py`` imports ``AutoModelForVision2Seq``, which the pinned ``transformers==4.56.2`` has and
  the installed transformers 5.5 does not.

So this arm runs the reference's *own* ``utils.get_adam`` / ``utils.get_scheduler`` /
``dataset.IterableStreamDataSet`` / ``dataset_wrapper`` unmodified from ``baseline/_ref`` and
re-creates the ~50 lines of ``training.py``'s loop around them with stock components only -- no
hypha_b200 model, kernel, optimizer or engine is imported on this path:

    model      transformers ``LlamaForCausalLM`` built from the Llama-3-8B config (random init; the
               reference does ``AutoModelForCausalLM.from_pretrained`` -- model.py:47-56 -- there are no
               checkpoints offline), SDPA attention, bf16 parameters
    optimizer  ``torch.optim.AdamW(params, lr)`` through the reference's ``get_adam`` (utils.py:56-64)
    schedule   the reference's ``get_scheduler`` (constant by default, utils.py:90-102)
    data       safetensors slice file -> reference ``IterableStreamDataSet`` -> ``DataLoader`` ->
               host->device copy per step (what ``accelerator.prepare`` does with
               ``distributed_type: NO``, test.yaml:3)
    step       ``zero_grad`` / ``model(**batch)`` / ``backward`` / ``optimizer.step`` /
               ``scheduler.step`` / ``loss.detach().cpu()`` every step (training.py:106-115)
    outer      pseudo-gradient ``theta - theta_prev`` (utils.py:118-123), mean over workers with
               ``torch.distributed.all_reduce`` (NCCL), unfused Nesterov ``m = mu*m + g``,
               ``u = lr*(mu*m + g)``, ``theta = theta_prev + u``
               (crates/worker/src/executor/parameter_server.rs:386-446, utils.py:105-115), all on the
               GPU -- the reference does this through CPU tensors and three disk round trips per round,
               so this arm flatters it

The reference trains in fp32 (``mixed_precision: "no"``); Llama-3-8B with fp32 AdamW does not fit one
B200 at the benchmark's 2x4096 tokens, so the arm runs bf16 parameters (the precision BASELINE.json
names), again in the reference's favour.  Gradient checkpointing is switched on only when the plain
plan runs out of memory.
"""
from __future__ import annotations

import json
import os
import statistics
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(REPO, "baseline", "_ref")

WHY_NOT_UNMODIFIED = (
    "Rust/libp2p workspace (no cargo offline); its Python executor installs into baseline/_ref but "
    "training.py needs accelerate==1.10.1 (absent) and model.py needs transformers==4.56.2 "
    "(AutoModelForVision2Seq is gone in the installed 5.5)")

LLAMA_DIMS = {
    "llama3-8b": dict(vocab_size=128256, hidden_size=4096, intermediate_size=14336, num_hidden_layers=32,
                      num_attention_heads=32, num_key_value_heads=8, head_dim=128),
    "llama3-1b": dict(vocab_size=128256, hidden_size=2048, intermediate_size=8192, num_hidden_layers=16,
                      num_attention_heads=32, num_key_value_heads=8, head_dim=64),
    "llama-small": dict(vocab_size=32000, hidden_size=1024, intermediate_size=4096, num_hidden_layers=4,
                        num_attention_heads=8, num_key_value_heads=4, head_dim=128),
    "llama-tiny": dict(vocab_size=1024, hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                       num_attention_heads=4, num_key_value_heads=2, head_dim=64),
}


def _reference_modules():
    """The reference's own helper modules, unmodified, if the install into baseline/_ref exists."""
    if not os.path.isdir(os.path.join(REF_DIR, "hypha", "accelerate_executor")):
        return None, None
    sys.path.insert(0, REF_DIR)
    try:
        import hypha.accelerate_executor.dataset as ref_dataset
        import hypha.accelerate_executor.utils as ref_utils

        return ref_utils, ref_dataset
    except Exception as exc:                                  # pragma: no cover - environment dependent
        print(f"[reference arm] baseline/_ref present but not importable: {exc}", file=sys.stderr)
        return None, None
    finally:
        sys.path.remove(REF_DIR)


def build_model(name: str, seq_len: int, device, checkpointing: bool):
    import torch
    from transformers import LlamaConfig, LlamaForCausalLM

    dims = LLAMA_DIMS[name]
    cfg = LlamaConfig(rope_theta=500000.0, rms_norm_eps=1e-5, max_position_embeddings=max(seq_len, 8192),
                      tie_word_embeddings=False, attn_implementation="sdpa", use_cache=False, **dims)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device(device):
            model = LlamaForCausalLM(cfg)
    finally:
        torch.set_default_dtype(prev)
    model.train()
    if checkpointing:
        model.gradient_checkpointing_enable()
    return model, cfg


class StockOuter:
    """DiLoCo outer step out of stock parts: all_reduce on the pseudo-gradient + unfused Nesterov."""

    def __init__(self, model, lr: float = 0.7, momentum: float = 0.9):
        import torch

        self.params = [p for p in model.parameters()]
        self.prev = [p.detach().clone() for p in self.params]
        self.mom = [torch.zeros_like(p) for p in self.params]
        self.lr, self.mu, self.round = lr, momentum, 0

    @staticmethod
    def _world():
        import torch.distributed as dist

        return dist.get_world_size() if dist.is_initialized() else 1

    def step(self):
        import torch
        import torch.distributed as dist

        with torch.no_grad():
            delta = torch._foreach_sub([p.detach() for p in self.params], self.prev)     # theta - theta_prev
            world = self._world()
            if world > 1:
                for d in delta:
                    dist.all_reduce(d)
                torch._foreach_div_(delta, world)
            if self.round == 0:
                for m, d in zip(self.mom, delta):
                    m.copy_(d)                                                            # m0 = g0
            else:
                torch._foreach_mul_(self.mom, self.mu)
                torch._foreach_add_(self.mom, delta)                                      # m = mu*m + g
            upd = torch._foreach_mul(self.mom, self.mu)
            torch._foreach_add_(upd, delta)
            torch._foreach_mul_(upd, self.lr)                                             # u = lr*(mu*m + g)
            torch._foreach_add_(self.prev, upd)                                           # theta_prev + u
            for p, q in zip(self.params, self.prev):
                p.copy_(q)
            self.round += 1


def write_slice(path: str, vocab: int, samples: int, seq_len: int, seed: int) -> None:
    """One dataset slice in the reference's on-disk format: safetensors with one row per sample."""
    import torch
    from safetensors.torch import save_file

    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, vocab, (samples, seq_len), generator=g)
    save_file({"input_ids": ids, "labels": ids.clone()}, path)


def run(args, ClockSampler) -> int:
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    on_gpu = torch.cuda.is_available()
    if not on_gpu and os.environ.get("HYPHA_BENCH_ALLOW_CPU") != "1":
        print(json.dumps({"impl": args.impl, "unavailable": "no CUDA device visible"}))
        return 0
    if on_gpu:
        torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank) if on_gpu else torch.device("cpu")     # cpu: plumbing test only
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if on_gpu:
            dist.init_process_group("nccl", device_id=dev)
        else:
            dist.init_process_group("gloo")

    class _WallEvent:                                  # stands in for a CUDA event on the CPU plumbing path
        def __init__(self, enable_timing=True):
            self.t = 0.0

        def record(self):
            self.t = time.perf_counter()

        def elapsed_time(self, other):
            return (other.t - self.t) * 1e3

    Event = torch.cuda.Event if on_gpu else _WallEvent

    def dev_sync():
        if on_gpu:
            torch.cuda.synchronize()
    ref_utils, ref_dataset = _reference_modules()
    B, S, H = args.micro_batch, args.seq_len, args.inner_steps
    lr = 1e-4

    def make(checkpointing: bool):
        model, cfg = build_model(args.model, S, dev, checkpointing)
        if ref_utils is not None:
            opt = ref_utils.get_adam({"learning-rate": lr}, model.parameters())        # reference utils.py:56-64
            sched = ref_utils.get_scheduler(None, opt)                                   # utils.py:90-102
        else:
            from transformers.optimization import get_constant_schedule

            opt = torch.optim.AdamW(model.parameters(), lr=lr)
            sched = get_constant_schedule(opt)
        return model, cfg, opt, sched, StockOuter(model)

    def one_step(model, opt, sched, batch, read_loss: bool):
        opt.zero_grad()
        outputs = model(**batch)
        loss = outputs if isinstance(outputs, torch.Tensor) else outputs["loss"]
        loss.backward()
        opt.step()
        sched.step()
        return loss.detach().cpu().numpy() if read_loss else loss.detach()

    g0 = torch.Generator().manual_seed(1000 + rank)
    ids = [torch.randint(0, LLAMA_DIMS[args.model]["vocab_size"], (B, S), generator=g0) for _ in range(4)]
    dev_pool = [{"input_ids": t.to(dev), "labels": t.to(dev)} for t in ids]

    # ---- memory plan: plain first, gradient checkpointing only if that does not fit --------------------
    state = None
    used_ckpt = False
    for ckpt in (False, True):
        ok = 1
        try:
            state = make(ckpt)
            for i in range(2):
                one_step(state[0], state[2], state[3], dev_pool[i % 4], read_loss=True)
            state[4].step()
            dev_sync()
        except torch.OutOfMemoryError:
            ok = 0
        if world > 1:
            t = torch.tensor([ok], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            ok = int(t.item())
        if ok:
            used_ckpt = ckpt
            break
        state = None
        import gc

        gc.collect()
        if on_gpu:
            torch.cuda.empty_cache()
        if rank == 0:
            print("[reference arm] plain plan does not fit; enabling gradient checkpointing", file=sys.stderr)
    if state is None:
        print(json.dumps({"impl": args.impl, "unavailable": "stock model does not fit one B200 even with "
                          "gradient checkpointing"}))
        return 0
    model, cfg, opt, sched, outer = state

    def barrier():
        if world > 1:
            dist.barrier()
        dev_sync()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- data path of the reference: slice file -> IterableStreamDataSet -> DataLoader ----------------
    tmp = tempfile.mkdtemp(prefix="hypha_ref_slice_")
    slice_path = os.path.join(tmp, f"slice_{rank}.safetensors")
    write_slice(slice_path, LLAMA_DIMS[args.model]["vocab_size"], 4 * B, S, 2000 + rank)

    def slice_files():
        while True:
            yield slice_path

    if ref_dataset is not None:
        ds = ref_dataset.IterableStreamDataSet(slice_files(), ["input_ids", "labels"], [], None)   # dataset.py:9-33
        loader = ref_dataset.dataset_wrapper(torch.utils.data.DataLoader(ds, batch_size=B))
    else:
        from safetensors.torch import load_file

        def gen():
            while True:
                d = load_file(slice_path, device="cpu")
                for i in range(0, d["input_ids"].shape[0], B):
                    yield {k: v[i:i + B] for k, v in d.items()}

        loader = gen()
    h2d = {"bytes": 0}

    def next_batch():
        b = next(loader)
        h2d["bytes"] += sum(v.numel() * v.element_size() for v in b.values())
        return {k: v.to(dev, non_blocking=True) for k, v in b.items()}               # accelerate's send_to_device

    tokens_per_step = B * S * world
    steps, warm = args.steps, max(args.warmup, 3)
    gstep = {"n": 0}

    def run_steps(n: int, e2e: bool) -> int:
        did = 0
        for i in range(n):
            if e2e:
                one_step(model, opt, sched, next_batch(), read_loss=True)            # loss.cpu() every step
            else:
                one_step(model, opt, sched, dev_pool[i % 4], read_loss=False)
            gstep["n"] += 1
            if gstep["n"] % H == 0:
                outer.step()
                did += 1
        if did == 0:
            outer.step()                  # same accounting as the product arm: >= 1 outer step per region
            did = 1
        return did

    for i in range(warm):
        one_step(model, opt, sched, next_batch(), read_loss=True)
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = Event(enable_timing=True), Event(enable_timing=True)
    barrier()
    ev0.record()
    outer_in_region = run_steps(steps, e2e=False)
    ev1.record()
    barrier()
    clocks = sampler.stop()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    value = tokens_per_step * steps / (ms_total / 1e3)

    h2d["bytes"] = 0
    barrier()
    t0 = time.perf_counter()
    run_steps(steps, e2e=True)
    barrier()
    wall = max_over_ranks(time.perf_counter() - t0)
    e2e = {"value": tokens_per_step * steps / wall, "unit": "tokens/s",
           "h2d_bytes_per_step": h2d["bytes"] // steps, "d2h_bytes_per_step": 4,
           "ms_per_step": wall * 1e3 / steps}

    outer_ms = None
    if world > 1:
        outer.step()
        barrier()
        o0, o1 = Event(enable_timing=True), Event(enable_timing=True)
        o0.record()
        outer.step()
        outer.step()
        o1.record()
        barrier()
        outer_ms = max_over_ranks(o0.elapsed_time(o1)) / 2

    if rank == 0:
        n_params = sum(p.numel() for p in model.parameters())
        print(json.dumps({
            "metric": "tokens/sec (whole box, device-timed, max over ranks) Llama-3-8B DiLoCo H=100",
            "value": value, "unit": "tokens/s", "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic tokens, random-init weights", "impl": args.impl,
            "reference_class": "stock-recreation",
            "reference_code_used": ("baseline/_ref: utils.get_adam, utils.get_scheduler, dataset.IterableStreamDataSet, "
                                    "dataset.dataset_wrapper (unmodified)") if ref_utils is not None else "none (baseline/_ref absent)",
            "why_not_unmodified": WHY_NOT_UNMODIFIED,
            "config": {
                "model": f"transformers.LlamaForCausalLM {args.model}", "params": n_params,
                "global_batch": B * world, "seq_len": S, "tokens_per_step": tokens_per_step,
                "parallelism": f"diloco-dp{world}", "inner_steps_H": H,
                "outer_steps_in_timed_region": outer_in_region,
                "outer_backend": "torch.distributed.all_reduce (NCCL) + unfused Nesterov",
                "optimizer": "torch.optim.AdamW (reference get_adam), bf16 parameters and moments",
                "attention": "torch SDPA", "gradient_checkpointing": used_ckpt,
                "l2": "no flush needed: >16 GB of weights/activations streamed per step >> 126 MB L2",
            },
            "outer_step_ms": outer_ms, "clocks": clocks, "e2e": e2e,
            "gpu_launches": 0,
        }))
    if world > 1:
        dist.destroy_process_group()
    try:
        os.remove(slice_path)
        os.rmdir(tmp)
    except OSError:
        pass
    return 0
