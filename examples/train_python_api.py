"""`DiLoCoTrainer` without the control plane: every GPU of one box runs H local AdamW steps on the sm_100a kernels, then
the fused peer-memory outer step synchronises them (no NCCL kernel on that path; the process group only bootstraps the
peer-memory handles).  Synthetic tokens; swap `batches()` for your own loader.

    python examples/train_python_api.py --model llama-small --steps 40                       # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \\
        examples/train_python_api.py --model llama3-8b --inner-steps 100 --steps 400            # one box

Writes a full-state checkpoint per rank (`--checkpoint-dir`; resume by passing the same directory again) and, on rank
0, a Hugging Face model directory (`--export`).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama-small", help="a preset of hypha_b200.models.llama.PRESETS")
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--inner-steps", type=int, default=10, help="H: local steps between outer steps")
    ap.add_argument("--micro-batch", type=int, default=2)
    ap.add_argument("--seq-len", type=int, default=1024)
    ap.add_argument("--lr", type=float, default=3e-4)
    ap.add_argument("--checkpoint-dir", default=None)
    ap.add_argument("--export", default=None)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from safetensors.torch import save_file

    from hypha_b200.models.llama import get_config
    from hypha_b200.parallel.diloco import AdamWConfig, DiLoCoConfig
    from hypha_b200.parallel.peer_memory import PeerGroup
    from hypha_b200.trainer import DiLoCoTrainer

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    group = PeerGroup.from_torch_distributed(local) if world > 1 else None

    cfg = get_config(args.model, max_seq_len=max(args.seq_len, 8))
    schedule = {"type": "cosine-with-warmup", "warmup_steps": max(1, args.steps // 20), "training_steps": args.steps}
    trainer = DiLoCoTrainer.create(cfg, adamw=AdamWConfig(lr=args.lr), diloco=DiLoCoConfig(inner_steps=args.inner_steps),
                                   group=group, device=local, seed=0, scheduler=schedule)
    ckpt = os.path.join(args.checkpoint_dir, f"rank{rank}.safetensors") if args.checkpoint_dir else None
    if ckpt and os.path.exists(ckpt):
        trainer.load_checkpoint(ckpt)
        print(f"rank {rank}: resumed at step {trainer.global_step}", flush=True)

    def batches():
        g = torch.Generator().manual_seed(1000 + rank)                 # every worker sees its own data
        while True:
            yield torch.randint(0, cfg.vocab_size, (args.micro_batch, args.seq_len), generator=g).pin_memory()

    data = batches()
    while trainer.global_step < args.steps:
        trainer.step(next(data))                                       # labels default to the next token
        if trainer.global_step % args.inner_steps == 0 and rank == 0:  # one 4-byte read per round, not per step
            print(json.dumps({"step": trainer.global_step, "loss": round(trainer.read_loss(), 4),
                              "outer_backend": trainer.outer.backend}), flush=True)
    if ckpt:
        os.makedirs(args.checkpoint_dir, exist_ok=True)
        trainer.save_checkpoint(ckpt)
    if args.export and rank == 0:
        os.makedirs(args.export, exist_ok=True)
        cfg.to_hf_json(os.path.join(args.export, "config.json"))
        save_file({k: v.contiguous().cpu() for k, v in trainer.model.to_hf_state_dict().items()},
                  os.path.join(args.export, "model.safetensors"), metadata={"format": "pt"})
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
