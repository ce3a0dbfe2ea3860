"""DiLoCoTrainer: the public training API of hypha_b200 on one B200 (one rank).

    trainer = DiLoCoTrainer.create("llama3-8b", group=PeerGroup.from_torch_distributed())
    loss = trainer.step(input_ids_pinned_cpu, labels_pinned_cpu)   # one inner step (+ outer every H)

A step = H2D copy of the batch on a side stream, forward + backward on the
sm_100a kernels, ONE fused flat-AdamW launch, and -- every ``inner_steps`` steps --
the fused NVLink outer step.  The loss stays on the device; ``loss.item()`` (or
``trainer.read_loss()``) is the only device->host sync the caller pays, which
removes the reference's per-step ``loss.cpu()`` + HTTP status round trip
(training.py:115-125).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from .models.llama import LlamaConfig, LlamaForCausalLM, get_config
from .parallel.diloco import AdamWConfig, DiLoCoConfig, DiLoCoOuter, FlatAdamW, make_lr_lambda
from .parallel.peer_memory import PeerGroup
from .telemetry import DeviceTimer, nvtx_range


@dataclass
class StepStats:
    step: int
    round: int
    did_outer: bool


class DiLoCoTrainer:
    def __init__(self, model: LlamaForCausalLM, adamw: AdamWConfig, diloco: DiLoCoConfig,
                 group: Optional[PeerGroup] = None, param_buf=None, grad_buf=None,
                 scheduler: Optional[dict] = None, process_group=None):
        self.model = model
        self.device = model.flat.data.device
        self.opt = FlatAdamW(model.flat.data, model.flat.grad, adamw, make_lr_lambda(scheduler))
        self.outer = DiLoCoOuter(self.opt, diloco, group, param_buf, grad_buf, process_group)
        self.diloco = diloco
        self.global_step = 0
        self._h2d_stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        # optimizer-in-backward: each layer's AdamW slice runs on a side stream as soon as that layer's
        # gradients exist, hidden under the remaining backward GEMMs (HBM-bound vs tensor-bound work)
        self.overlap_optimizer = os.environ.get("HYPHA_OVERLAP_OPT", "1") == "1" and self.device.type == "cuda"
        self._opt_stream = torch.cuda.Stream(device=self.device) if self.overlap_optimizer else None
        self._opt_done: list = []
        if self.overlap_optimizer:
            model.grad_ready_hook = self._on_grad_ready
        self.last_loss: Optional[torch.Tensor] = None
        self._main_grad = None              # fp32 gradient sum, created by the first inner_step_accumulated("fp32")
        self.h2d_bytes = 0
        self.time_outer = os.environ.get("HYPHA_TIME_OUTER", "0") == "1"
        self.last_outer_ms: Optional[float] = None

    # -- construction -------------------------------------------------------------------
    @classmethod
    def create(cls, model: str | LlamaConfig = "llama3-8b", *, adamw: Optional[AdamWConfig] = None,
               diloco: Optional[DiLoCoConfig] = None, group: Optional[PeerGroup] = None,
               device: Optional[int] = None, seed: int = 0, scheduler: Optional[dict] = None,
               process_group=None, **model_overrides) -> "DiLoCoTrainer":
        cfg = get_config(model, **model_overrides) if isinstance(model, str) else model
        adamw = adamw or AdamWConfig()
        diloco = diloco or DiLoCoConfig()
        dev = torch.cuda.current_device() if device is None else device
        torch.cuda.set_device(dev)
        param_buf = grad_buf = None
        # peer memory (and with it the fused outer step) ends at the NVSwitch box: a larger team trains with the
        # sharded outer step over the process group (`DiLoCoOuter._step_sharded`)
        want_fused = group is not None and group.world > 1 and (
            diloco.backend == "fused" or (diloco.backend == "auto" and group.world <= 8))
        if want_fused:
            holders = {}
            nvls_mode = os.environ.get("HYPHA_NVLS", "1")      # "0" off | "1" after a local pre-flight | "force"
            # a joining process allocates alone (PeerGroup.joining): no collective NVLS set-up
            use_nvls = nvls_mode != "0" and group.world >= 4 and not getattr(group, "deferred", False)

            def sym_alloc(name, n, dt):
                """NVLS (multicast-bound VMM) buffer when the fabric supports it, else CUDA-IPC.  All
                ranks take the same branch: the decision is agreed through the control plane."""
                nonlocal use_nvls
                if use_nvls:
                    from .parallel.nvls import NvlsUnavailable, alloc_nvls, multicast_supported, nvls_usable

                    # every rank probes locally (no communication), then ONE aligned exchange decides
                    mine = multicast_supported(dev) if nvls_mode == "force" else nvls_usable(dev, group.world)
                    ok = all(group._exchange(bool(mine)))
                    if ok:
                        try:
                            return alloc_nvls(group, name, n, dt)
                        except NvlsUnavailable as exc:
                            # alloc_nvls reports a failure to EVERY rank at the same exchange and undoes its
                            # mappings, so all ranks arrive here together and the group is still aligned
                            if nvls_mode == "force":
                                raise RuntimeError(f"NVLS set-up failed ({exc}); rerun with HYPHA_NVLS=0") from exc
                            if group.rank == 0:
                                print(f"[trainer] NVLS set-up failed ({exc}); using CUDA-IPC peer buffers", flush=True)
                    use_nvls = False
                return group.alloc(name, n, dt)

            def alloc(n, dt):
                holders["p"] = sym_alloc("model.params", n, dt)
                return holders["p"].local

            def galloc(n, dt):
                holders["g"] = sym_alloc("model.grads", n, dt)
                return holders["g"].local

            net = LlamaForCausalLM(cfg, device=f"cuda:{dev}", alloc=alloc, grad_alloc=galloc, seed=seed)
            param_buf, grad_buf = holders["p"], holders["g"]
        else:
            net = LlamaForCausalLM(cfg, device=f"cuda:{dev}", seed=seed)
        return cls(net, adamw, diloco, group, param_buf, grad_buf, scheduler, process_group)

    # -- one inner step ---------------------------------------------------------------------
    def _to_device(self, t: torch.Tensor) -> torch.Tensor:
        if t.device == self.device:
            return t
        self.h2d_bytes += t.numel() * t.element_size()
        with torch.cuda.stream(self._h2d_stream):
            d = t.to(self.device, non_blocking=True)
        torch.cuda.current_stream().wait_stream(self._h2d_stream)
        d.record_stream(torch.cuda.current_stream())
        return d

    def _on_grad_ready(self, begin: int, end: int) -> None:
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self._opt_stream.wait_event(ev)
        with torch.cuda.stream(self._opt_stream):
            self.opt.step_range(begin, end)
        self._opt_done.append((begin, end))

    def inner_step(self, input_ids: torch.Tensor, labels: torch.Tensor, ignore_index: Optional[int] = None,
                   n_valid: Optional[int] = None) -> torch.Tensor:
        """``labels[b, t]`` is the target of position ``t`` (already shifted); ``ignore_index`` / ``n_valid``: masked
        targets and the number of unmasked ones (see `LlamaForCausalLM.forward`)."""
        # NVTX ranges (hypha.h2d / fwd / bwd / opt / outer) mark the phases of a step in ncu / nsys timelines
        with nvtx_range("hypha.h2d"):
            ids = self._to_device(input_ids)
            lab = self._to_device(labels)
        self.model.set_grad_accumulate(False)
        if not self.overlap_optimizer:
            with nvtx_range("hypha.fwd"):
                loss = self.model(ids, lab, ignore_index=ignore_index, n_valid=n_valid)
            with nvtx_range("hypha.bwd"):
                loss.backward()
            with nvtx_range("hypha.opt"):
                self.opt.step()
        else:
            self.opt.begin_step()
            self._opt_done.clear()
            with nvtx_range("hypha.fwd"):
                loss = self.model(ids, lab, ignore_index=ignore_index, n_valid=n_valid)
            with nvtx_range("hypha.bwd+opt"):          # each layer's AdamW slice runs under the remaining backward
                loss.backward()
            # whatever backward did not cover (the embedding table, or everything if no hook fired)
            done = sorted(self._opt_done)
            cur, n = 0, self.opt.params.numel()
            gaps = []
            for b, e in done:
                if b > cur:
                    gaps.append((cur, b))
                cur = max(cur, e)
            if cur < n:
                gaps.append((cur, n))
            for b, e in gaps:
                self._on_grad_ready(b, e)
            torch.cuda.current_stream().wait_stream(self._opt_stream)
            self.opt.finish_step()
        self.last_loss = loss.detach()
        return self.last_loss

    def inner_step_accumulated(self, micro_batches, main_grad: str = "fp32", ignore_index: Optional[int] = None
                               ) -> torch.Tensor:
        """One inner step whose gradient is the mean over several micro-batches (``[(input_ids, labels), ...]`` or
        ``[(input_ids, labels, n_valid), ...]``, labels already shifted as in `inner_step`).  ``main_grad``:

        * ``"fp32"`` -- each micro-batch writes its bf16 gradients in overwrite mode and they are summed in an fp32
          buffer (`parallel.main_grad.MainGrad`), rounded to bf16 once before the fused AdamW;
        * ``"bf16"`` -- the wgrad GEMM epilogues add into the bf16 gradient buffer (no extra memory or traffic, one
          rounding per micro-batch), the 1/k goes into the AdamW launch as ``grad_scale``.

        The optimizer cannot start before the last micro-batch, so this path does not use the optimizer-in-backward
        overlap.  Returns the mean loss (on the device)."""
        if main_grad not in ("fp32", "bf16"):
            raise ValueError(f"main_grad must be 'fp32' or 'bf16', not {main_grad!r}")
        micro_batches = list(micro_batches)
        if not micro_batches:
            raise ValueError("inner_step_accumulated needs at least one micro-batch")
        k = len(micro_batches)
        if main_grad == "fp32" and self._main_grad is None:
            from .parallel.main_grad import MainGrad

            self._main_grad = MainGrad(self.model.flat.grad)
        hook, self.model.grad_ready_hook = self.model.grad_ready_hook, None
        total = None
        try:
            for i, mb in enumerate(micro_batches):
                input_ids, labels = mb[0], mb[1]
                n_valid = mb[2] if len(mb) > 2 else None       # known count of unmasked targets: no host sync
                with nvtx_range("hypha.h2d"):
                    ids, lab = self._to_device(input_ids), self._to_device(labels)
                self.model.set_grad_accumulate(main_grad == "bf16" and i > 0)
                with nvtx_range("hypha.fwd"):
                    loss = self.model(ids, lab, ignore_index=ignore_index, n_valid=n_valid)
                with nvtx_range("hypha.bwd"):
                    loss.backward()
                if main_grad == "fp32":
                    self._main_grad.add()
                total = loss.detach() if total is None else total + loss.detach()
        finally:
            self.model.grad_ready_hook = hook
            self.model.set_grad_accumulate(False)
        with nvtx_range("hypha.opt"):
            if main_grad == "fp32":
                self._main_grad.finalize()
                self.opt.step()
            else:
                self.opt.step(grad_scale=1.0 / k)
        self.last_loss = total / k
        return self.last_loss

    def outer_step(self) -> None:
        """One DiLoCo outer step; its device time (CUDA events on the launching stream) is kept in
        ``last_outer_ms`` when ``time_outer`` is set (costs one event synchronisation)."""
        with nvtx_range("hypha.outer"):
            if self.time_outer and self.device.type == "cuda":
                with DeviceTimer() as t:
                    self.outer.step()
                self.last_outer_ms = t.ms
            else:
                self.outer.step()

    def step(self, input_ids: torch.Tensor, labels: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One DiLoCo inner step; runs the outer step when ``inner_steps`` are complete."""
        if labels is None:
            labels = torch.roll(input_ids, -1, dims=-1)
        loss = self.inner_step(input_ids, labels)
        self.global_step += 1
        if self.global_step % self.diloco.inner_steps == 0:
            self.outer_step()
        return loss

    def read_loss(self) -> float:
        """Device->host read of the last loss (4 bytes)."""
        return float(self.last_loss.item())

    # -- checkpoint (absent in the reference; needed for rejoin) -------------------
    def state_dict(self) -> Dict[str, torch.Tensor]:
        sd = {"params": self.model.flat.data, "global_step": torch.tensor(self.global_step)}
        sd.update({f"opt.{k}": v for k, v in self.opt.state_dict().items()})
        sd.update({f"outer.{k}": v for k, v in self.outer.state_dict().items()})
        return sd

    def save_checkpoint(self, path: str) -> None:
        from safetensors.torch import save_file

        tmp = path + ".tmp"
        save_file({k: v.detach().contiguous().cpu() for k, v in self.state_dict().items()}, tmp)
        os.replace(tmp, path)

    def load_checkpoint(self, path: str) -> None:
        from safetensors.torch import load_file

        sd = load_file(path, device=str(self.device))
        self.model.flat.data.copy_(sd["params"])
        self.global_step = int(sd["global_step"])
        self.opt.load_state_dict({k[4:]: v for k, v in sd.items() if k.startswith("opt.")})
        self.outer.load_state_dict({k[6:]: v for k, v in sd.items() if k.startswith("outer.")})
