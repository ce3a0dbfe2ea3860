"""DiLoCo training executor: ``python -m hypha_b200.executor.training --socket S --work-dir W --job J``.

Same contract as the reference executor (executors/accelerate/.../training.py:28-162): talk to
the worker's job bridge over a Unix socket, fetch model + data, run inner AdamW steps while
reporting ``{"status": {"batch_size": n}}`` after every batch, and when the scheduler answers
``ScheduleUpdate{counter}`` count down, ship the pseudo-gradient ``theta - theta_prev`` to the
aggregator, wait for the outer update, apply ``theta = theta_prev + update`` and acknowledge.

What is different:
  * theta_prev is kept in memory (device) instead of a safetensors file that is re-read twice per
    round, and the model never leaves the device (Q11);
  * the loss is accumulated on the device and read back once per round (Q10);
  * ``round_time`` is dropped (Q2); models are local architectures (Q12); WSD reads
    ``decay_steps`` (Q4);
  * on CUDA the Llama family trains with the sm_100a kernels + fused flat AdamW.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from typing import Any, Dict, Iterator, List, Optional

import torch

from .api import Session
from .dataset import IterableStreamDataSet, forever
from .model import _LlamaAdapter, get_model, get_preprocessor

FETCH_PATH = "artifacts"


def prepare_files(config: dict, session: Session) -> None:
    session.fetch(config["model"]["artifact"])
    if config.get("preprocessor"):
        session.fetch(config["preprocessor"]["artifact"])


def fetch_data(session: Session, data: Any, work_dir: str) -> Iterator[str]:
    while True:
        got = session.fetch(data)
        yield os.path.join(work_dir, got[0]["path"])


class InnerLoop:
    """Model + inner optimizer + the bookkeeping DiLoCo needs (snapshot / delta / merge)."""

    def __init__(self, model: torch.nn.Module, opt_cfg: dict, sched_cfg: Optional[dict], device: str):
        from ..parallel.diloco import AdamWConfig, FlatAdamW, make_lr_lambda

        self.model, self.device = model, device
        self.flat = isinstance(model, _LlamaAdapter) and device.startswith("cuda")
        lr = float(opt_cfg["learning-rate"])
        betas = tuple(opt_cfg["betas"]) if opt_cfg.get("betas") else (0.9, 0.999)
        eps = float(opt_cfg["epsilon"]) if opt_cfg.get("epsilon") else 1e-8
        lam = make_lr_lambda(sched_cfg)
        if self.flat:
            net = model.net
            self.opt = FlatAdamW(net.flat.data, net.flat.grad, AdamWConfig(lr=lr, betas=betas, eps=eps,
                                                                          state_dtype="fp32"), lam)
            self.prev = self.opt.master.clone()
        else:
            self.opt = torch.optim.AdamW(model.parameters(), lr=lr, betas=betas, eps=eps)
            self.sched = torch.optim.lr_scheduler.LambdaLR(self.opt, lam)
            self.prev = {k: v.detach().clone() for k, v in model.state_dict().items()}

    def step(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        batch = {k: v.to(self.device, non_blocking=True) for k, v in batch.items()}
        if not self.flat:
            self.opt.zero_grad(set_to_none=True)
        out = self.model(**batch)
        loss = out if isinstance(out, torch.Tensor) else out["loss"]
        loss.backward()
        self.opt.step()
        if not self.flat:
            self.sched.step()
        return loss.detach()

    def delta_state_dict(self) -> Dict[str, torch.Tensor]:
        """theta - theta_prev per tensor, fp32 on CPU (utils.py:118-123)."""
        if self.flat:
            d = (self.opt.master - self.prev)
            net = self.model.net
            return {n: net.flat.view(n, of=d).detach().float().cpu().contiguous() for n in net.flat.layout}
        sd = self.model.state_dict()
        return {k: (v.detach().float() - self.prev[k].float()).cpu().contiguous() for k, v in sd.items()
                if v.dtype.is_floating_point}

    def apply_update(self, update: Dict[str, torch.Tensor]) -> None:
        """theta = theta_prev + update; becomes the new theta_prev (utils.py:105-115)."""
        if self.flat:
            net = self.model.net
            for n in net.flat.layout:
                net.flat.view(n, of=self.prev).add_(update[n].to(self.prev.device).view(net.flat.layout[n][1]))
            self.opt.master.copy_(self.prev)
            net.flat.data.copy_(self.prev)
            return
        sd = self.model.state_dict()
        with torch.no_grad():
            for k, v in sd.items():
                if k in update:
                    new = self.prev[k].float() + update[k].to(v.device)
                    v.copy_(new.to(v.dtype))
                    self.prev[k] = v.detach().clone()


def main(socket_path: str, work_dir: str, job_json: str) -> None:
    from safetensors.torch import load_file, save_file

    device = "cuda" if torch.cuda.is_available() else "cpu"
    with Session(socket_path) as session:
        job = json.loads(job_json)
        executor = job["executor"]
        assert executor["class"] == "train"
        config = executor["config"]
        prepare_files(config, session)
        local = os.path.join(work_dir, FETCH_PATH)
        model = get_model(local, config["model"]["task"], device)
        loop = InnerLoop(model, config["optimizer"], config.get("scheduler"), device)
        pre_cfg = config.get("preprocessor")
        loader = torch.utils.data.DataLoader(
            IterableStreamDataSet(fetch_data(session, config["data"], work_dir), config["model"]["input-names"],
                                  pre_cfg["input-names"] if pre_cfg else [], get_preprocessor(pre_cfg, local)),
            batch_size=config["batch_size"])
        batches = forever(loader)
        round_no = 1
        with session.receive(config["results"], "incoming") as receiver:
            updates = iter(receiver)
            while True:
                loss_sum = torch.zeros((), device=device)
                steps, counter = 0, -1
                while counter != 0:
                    batch = next(batches)
                    loss_sum += loop.step(batch)
                    steps += 1
                    bs = int(next(iter(batch.values())).shape[0])
                    resp = session.send_status({"status": {"batch_size": bs}})
                    if resp["type"] == "ScheduleUpdate":
                        counter = resp["counter"]
                    elif resp["type"] == "Error":
                        raise RuntimeError("scheduler rejected status")
                    else:
                        counter -= 1
                session.send_status("update")
                name = f"{round_no}_local_gradients.safetensors"
                save_file(loop.delta_state_dict(), os.path.join(work_dir, name))
                session.send_resource(config["updates"], name)
                session.send_status({"metrics": {"round": round_no, "metrics": {"loss": float(loss_sum.item() / max(1, steps))}}})
                print(f"round {round_no}: {steps} inner steps, mean loss {float(loss_sum.item() / max(1, steps)):.4f}; "
                      f"waiting for the outer update", flush=True)
                event = next(updates)
                latest = event[-1] if isinstance(event, list) else event
                rel = (latest.get("parameters") or {}).get("path") if isinstance(latest.get("parameters"), dict) else latest.get("path")
                loop.apply_update(load_file(os.path.join(work_dir, rel)))
                os.unlink(os.path.join(work_dir, name))
                round_no += 1
                if session.send_status("update-received")["type"] == "Done":
                    break
        print(f"Finished training of {round_no - 1} DiLoCo update rounds", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--socket", required=True)
    ap.add_argument("--work-dir", required=True)
    ap.add_argument("--job", required=True)
    a = ap.parse_args()
    main(a.socket, a.work_dir, a.job)
