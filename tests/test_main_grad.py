"""fp32 main gradients for gradient accumulation (`parallel.main_grad`, `DiLoCoTrainer.inner_step_accumulated`)."""
import pytest
import torch

from hypha_b200.parallel.main_grad import MainGrad


def test_fp32_sum_rounds_once_where_a_bf16_sum_drifts():
    """256 micro-batch gradients of similar size: the running bf16 sum stops absorbing addends once it is ~2^8 times
    larger than one of them, the fp32 sum rounded once at the end stays within one bf16 ulp of the true mean."""
    torch.manual_seed(0)
    k, n = 256, 4096
    micro = (1.0 + 0.1 * torch.randn(k, n)).to(torch.bfloat16)
    truth = micro.double().mean(0)

    grads = torch.zeros(n, dtype=torch.bfloat16)
    mg = MainGrad(grads)
    running = torch.zeros(n, dtype=torch.bfloat16)
    for i in range(k):
        grads.copy_(micro[i])                 # what an overwrite-mode backward leaves in the gradient buffer
        mg.add()
        running += micro[i]                   # what the accumulate-mode GEMM epilogue does
    assert mg.finalize() == k and mg.count == 0
    err_fp32 = (grads.double() - truth).abs().max().item()
    err_bf16 = ((running.float() / k).double() - truth).abs().max().item()
    assert err_fp32 <= 2 ** -8 * 1.01          # half an ulp of bf16 at ~1.0
    assert err_bf16 > 4 * err_fp32             # the bf16 running sum is several times worse (and grows with k)


def test_single_micro_batch_is_the_identity_and_state_resets():
    g = torch.randn(64).to(torch.bfloat16)
    keep = g.clone()
    mg = MainGrad(g)
    mg.add()
    assert mg.finalize() == 1 and torch.equal(g, keep)
    with pytest.raises(RuntimeError):
        mg.finalize()
    g.fill_(2.0)
    mg.add()
    g.fill_(4.0)
    mg.add()
    mg.finalize()
    assert torch.all(g == 3.0)                 # the previous sum did not leak into this one
    mg.release()
    assert mg.acc is None


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_trainer_accumulated_step_uses_the_mean_gradient(mode):
    """`inner_step_accumulated` hands the optimizer the mean of the micro-batch gradients (fp32: in the gradient
    buffer; bf16: the sum there and 1/k as the AdamW grad_scale), switches the optimizer-in-backward hook off for the
    duration and leaves accumulate mode off afterwards."""
    from hypha_b200.models.llama import LlamaConfig, LlamaForCausalLM
    from hypha_b200.parallel.diloco import AdamWConfig, DiLoCoConfig
    from hypha_b200.trainer import DiLoCoTrainer

    cfg = LlamaConfig(vocab_size=64, hidden_size=32, intermediate_size=64, num_layers=2, num_heads=2, num_kv_heads=2,
                      max_seq_len=16)
    net = LlamaForCausalLM(cfg, device="cpu", seed=1)
    tr = DiLoCoTrainer(net, AdamWConfig(), DiLoCoConfig())
    seen = {}

    def fake_step(grad_scale=1.0):
        seen["grad"] = net.flat.grad.float() * grad_scale
    tr.opt.step = fake_step
    fired = []
    marker = lambda b, e: fired.append((b, e))
    net.grad_ready_hook = marker

    torch.manual_seed(0)
    batches = []
    for _ in range(3):
        ids = torch.randint(0, 64, (2, 16))
        batches.append((ids, torch.roll(ids, -1, dims=-1)))
    want, losses = torch.zeros(net.flat.numel), []
    net.grad_ready_hook = None
    for ids, lab in batches:
        net.set_grad_accumulate(False)
        loss = net(ids, lab)
        loss.backward()
        want += net.flat.grad.float() / 3
        losses.append(loss.item())
    net.grad_ready_hook = marker

    loss = tr.inner_step_accumulated(batches, main_grad=mode)
    assert abs(loss.item() - sum(losses) / 3) < 1e-5
    assert not fired and net.grad_ready_hook is marker
    assert all(not getattr(p, "grad_accumulate", False) for p in net.parameters())
    tol = 2e-3 if mode == "fp32" else 1e-2
    assert (seen["grad"] - want).abs().max().item() <= tol * want.abs().max().item() + 1e-6
    with pytest.raises(ValueError):
        tr.inner_step_accumulated(batches, main_grad="fp16")
    with pytest.raises(ValueError):
        tr.inner_step_accumulated([])


def test_split_batch_keeps_micro_batches_equal():
    from hypha_b200.executor.training import accumulation_of, split_batch

    b = {"input_ids": torch.arange(24).view(6, 4), "labels": torch.arange(24).view(6, 4)}
    parts = split_batch(b, 3)
    assert [p["input_ids"].shape[0] for p in parts] == [2, 2, 2]
    assert torch.equal(torch.cat([p["labels"] for p in parts]), b["labels"])
    assert len(split_batch(b, 4)) == 3            # 6 samples cannot make 4 equal parts: 3 of 2
    assert len(split_batch(b, 16)) == 6 and split_batch(b, 1) == [b]
    five = {"input_ids": torch.zeros(5, 4)}
    assert len(split_batch(five, 2)) == 1         # prime-sized tail batch: one micro-batch
    assert accumulation_of({}) == (1, "fp32")
    assert accumulation_of({"accumulation": {"micro_batches": 4, "main_grad": "bf16"}}) == (4, "bf16")
    with pytest.raises(ValueError):
        accumulation_of({"accumulation": {"micro_batches": 0}})


def test_inner_loop_accumulation_matches_the_full_batch_step():
    """Parameter-server / generic executor path (torch model + torch.optim.AdamW): one step on 4 micro-batches of 2
    equals one step on the batch of 8 (the loss is a mean over samples, so the mean of the micro-batch gradients is
    the batch gradient)."""
    from hypha_b200.executor.training import InnerLoop

    def make():
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 1))

        class Wrap(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.net = net

            def forward(self, x, y):
                return {"loss": ((self.net(x) - y) ** 2).mean()}
        return Wrap()

    opt = {"learning-rate": 1e-2}
    a, b = InnerLoop(make(), opt, None, "cpu"), InnerLoop(make(), opt, None, "cpu", (4, "fp32"))
    torch.manual_seed(1)
    batch = {"x": torch.randn(8, 8), "y": torch.randn(8, 1)}
    la, lb = a.step(batch), b.step(batch)
    assert abs(la.item() - lb.item()) < 1e-6
    for pa, pb in zip(a.model.parameters(), b.model.parameters()):
        assert torch.allclose(pa, pb, atol=1e-6)
    assert set(a.delta_state_dict()) == set(b.delta_state_dict())


def test_accumulation_travels_from_the_job_file_to_the_executor(tmp_path):
    import hypha_b200.config as C
    import hypha_b200.messages as M

    p = tmp_path / "s.toml"
    p.write_text('[scheduler.job.precision]\nmicro_batches = 4\nmain_grad = "bf16"\n')
    cfg = C.builder(C.SchedulerConfig).with_toml(str(p)).build()
    assert (cfg.scheduler.job.precision.micro_batches, cfg.scheduler.job.precision.main_grad) == (4, "bf16")
    for bad in ('micro_batches = 0\n', 'main_grad = "fp64"\n'):
        p.write_text('[scheduler.job.precision]\n' + bad)
        with pytest.raises(C.ConfigError) as exc:
            C.validate(C.builder(C.SchedulerConfig).with_toml(str(p)).build())
        assert exc.value.key.startswith("scheduler.job.precision.")
    wire = {"model": M.Model("causal-lm", M.Fetch.uri("https://example.org/m"), ["input_ids"]).to_wire(),
            "data": M.Fetch.scheduler("s", "d").to_wire(), "updates": M.Send.to_peers(["ps"]).to_wire(),
            "results": M.Receive.from_peers(["ps"]).to_wire(), "optimizer": M.Adam(1e-3).to_wire(), "batch_size": 8,
            "accumulation": {"micro_batches": 4, "main_grad": "bf16"}}
    t = M.TrainExecutorConfig.from_wire(wire)
    assert t.to_wire()["accumulation"] == {"micro_batches": 4, "main_grad": "bf16"}
    wire["accumulation"] = {"micro_batches": 1}
    assert "accumulation" not in M.TrainExecutorConfig.from_wire(wire).to_wire()     # reference job JSON unchanged
    wire["accumulation"] = {"micro_batches": -2}
    with pytest.raises(M.MessageError):
        M.TrainExecutorConfig.from_wire(wire)
