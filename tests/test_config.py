"""Layered config tests (model: crates/config/src/lib.rs:642-910, worker/src/config.rs:618-651)."""
import tomllib

import pytest

from hypha_b200 import config as C


def test_defaults_and_documented_toml_roundtrip(tmp_path):
    cfg = C.builder(C.WorkerConfig).build()
    assert cfg.resources.gpu == 16.0 and cfg.offer.strategy.type == "flexible" and cfg.work_dir == "/tmp"
    text = C.to_toml(cfg)
    assert "# capacity offered to schedulers" in text and "[[executors]]" in text
    parsed = tomllib.loads(text)
    assert parsed["executors"][0]["class"] == "train"
    p = tmp_path / "w.toml"
    p.write_text(text)
    again = C.builder(C.WorkerConfig).with_toml(str(p)).build()
    assert C.to_dict(again) == C.to_dict(cfg)


def test_precedence_toml_env_otel_cli(tmp_path):
    p = tmp_path / "s.toml"
    p.write_text('status_bridge = "file:1"\ngateway_addresses = ["tcp://1.2.3.4:1"]\ntelemetry_sampler = "always_off"\n')
    env = {"HYPHA_STATUS_BRIDGE": "env:2", "OTEL_TRACES_SAMPLER": "always_on", "OTEL_TRACES_SAMPLER_ARG": "0.25",
           "HYPHA_LISTEN_ADDRESSES": '["tcp://127.0.0.1:7"]', "UNRELATED": "x"}
    cfg = C.builder(C.SchedulerConfig).with_toml(str(p)).with_env(env).build()
    assert cfg.status_bridge == "env:2"                       # HYPHA_ beats file
    assert cfg.telemetry_sampler == "always_on"               # OTEL_ beats file (alias traces_sampler)
    assert cfg.telemetry_sample_ratio == 0.25
    assert cfg.listen_addresses == ["tcp://127.0.0.1:7"]
    assert cfg.gateway_addresses == ["tcp://1.2.3.4:1"]
    cfg2 = C.builder(C.SchedulerConfig).with_toml(str(p)).with_env(env).with_overrides({"status_bridge": "cli:3"}).build()
    assert cfg2.status_bridge == "cli:3"                      # CLI beats everything
    assert cfg2._meta.source["status_bridge"] == "cli"


def test_scheduler_job_defaults_match_reference():
    job = C.builder(C.SchedulerConfig).build().scheduler.job
    assert job.type == "diloco" and job.model.repository == "hypha-space/lenet"
    assert job.rounds.update_rounds == 100 and job.rounds.avg_samples_between_updates == 1200
    assert job.rounds.max_batch_size == 600
    assert job.inner_optimizer.learning_rate == 1e-3 and job.outer_optimizer.learning_rate == 0.7
    assert job.outer_optimizer.momentum == 0.9 and job.resources.num_workers == 2
    assert job.resources.worker.gpu == 10.0 and job.resources.worker_price.bid == 100.0
    d = C.to_dict(job)
    assert "learning-rate" in d["inner_optimizer"]           # kebab-case on the wire (messages :652-658)


def test_job_section_from_toml(tmp_path):
    p = tmp_path / "s.toml"
    p.write_text('''
[scheduler.job]
type = "diloco"
[scheduler.job.model]
repository = "local/llama3-8b"
filenames = ["config.json"]
type = "causal-lm"
input_names = ["input_ids", "labels"]
[scheduler.job.rounds]
avg_samples_between_updates = 6400
update_rounds = 3
[scheduler.job.inner_optimizer]
learning-rate = 0.0003
[scheduler.job.resources]
num_workers = 8
''')
    cfg = C.validate(C.builder(C.SchedulerConfig).with_toml(str(p)).build())
    job = cfg.scheduler.job
    assert job.model.type == "causal-lm" and job.resources.num_workers == 8
    assert job.inner_optimizer.learning_rate == 3e-4 and job.rounds.max_batch_size == 600


def test_unknown_key_and_type_errors_point_at_file(tmp_path):
    p = tmp_path / "w.toml"
    p.write_text('work_dir = "/tmp"\nbogus_key = 1\n')
    with pytest.raises(C.ConfigError) as e:
        C.builder(C.WorkerConfig).with_toml(str(p)).build()
    assert "bogus_key" in str(e.value) and "w.toml:2" in str(e.value)
    p.write_text('[resources]\ngpu = "many"\n')
    with pytest.raises(C.ConfigError) as e:
        C.builder(C.WorkerConfig).with_toml(str(p)).build()
    assert "resources.gpu" in str(e.value)


def test_gateway_in_excluded_cidr_is_rich_diagnostic(tmp_path):
    p = tmp_path / "w.toml"
    p.write_text('# worker\ngateway_addresses = ["tcp://10.1.2.3:4001"]\n')
    cfg = C.builder(C.WorkerConfig).with_toml(str(p)).build()
    with pytest.raises(C.ConfigError) as e:
        C.validate(cfg)
    msg = str(e.value)
    assert "10.0.0.0/8" in msg and "w.toml:2" in msg and "gateway_addresses =" in msg
    ok = C.builder(C.WorkerConfig).with_toml(str(p)).with_overrides({"exclude_cidr": []}).build()
    C.validate(ok)


def test_missing_file_and_bad_toml(tmp_path):
    with pytest.raises(C.ConfigError):
        C.builder(C.GatewayConfig).with_toml(str(tmp_path / "nope.toml")).build()
    bad = tmp_path / "bad.toml"
    bad.write_text("this is = = not toml")
    with pytest.raises(C.ConfigError):
        C.builder(C.GatewayConfig).with_toml(str(bad)).build()


def test_worker_validation_rules():
    cfg = C.builder(C.WorkerConfig).with_overrides({"offer": {"strategy": {"type": "greedy"}}}).build()
    with pytest.raises(C.ConfigError):
        C.validate(cfg)
    cfg = C.builder(C.WorkerConfig).with_overrides(
        {"executors": [{"class": "train", "name": "x", "runtime": "process"}]}).build()
    with pytest.raises(C.ConfigError):
        C.validate(cfg)


def test_tls_loader(tmp_path):
    cert = tmp_path / "c.pem"
    cert.write_bytes(b"-----BEGIN CERTIFICATE-----\n")
    cfg = C.builder(C.GatewayConfig).with_overrides({"cert_pem": str(cert)}).build()
    assert C.load_tls(cfg)["cert_pem"].startswith(b"-----BEGIN")
    cfg = C.builder(C.GatewayConfig).with_overrides({"key_pem": str(tmp_path / "missing.pem")}).build()
    with pytest.raises(C.ConfigError):
        C.load_tls(cfg)


def test_scheduler_job_aggregation_and_elastic_sections(tmp_path):
    """B200 extensions of [scheduler.job]: `aggregation.mode` is validated, `elastic` defaults on, both
    survive an init -> TOML -> load round trip."""
    from hypha_b200 import config as C

    cfg = C.builder(C.SchedulerConfig).with_overrides(
        {"scheduler": {"job": {"aggregation": {"mode": "collective", "rendezvous": "tcp://127.0.0.1:29400"},
                               "elastic": {"min_workers": 2}}}}).build()
    C.validate(cfg)
    path = tmp_path / "s.toml"
    path.write_text(C.to_toml(cfg))
    back = C.builder(C.SchedulerConfig).with_toml(str(path)).build()
    assert back.scheduler.job.aggregation.mode == "collective"
    assert back.scheduler.job.aggregation.rendezvous == "tcp://127.0.0.1:29400"
    assert back.scheduler.job.elastic.enabled is True and back.scheduler.job.elastic.min_workers == 2
    bad = C.builder(C.SchedulerConfig).with_overrides({"scheduler": {"job": {"aggregation": {"mode": "ring"}}}}).build()
    with pytest.raises(C.ConfigError) as ei:
        C.validate(bad)
    assert "unknown aggregation mode" in str(ei.value)
    bad2 = C.builder(C.SchedulerConfig).with_overrides(
        {"scheduler": {"job": {"aggregation": {"mode": "collective", "rendezvous": "udp://x"}}}}).build()
    with pytest.raises(C.ConfigError):
        C.validate(bad2)


def test_nested_env_keys_with_double_underscore():
    """`HYPHA_<KEY>` covers top-level keys like the reference; `__` reaches nested ones and JSON values
    still work.  Unknown top-level names are ignored (other tools also use the HYPHA_ prefix)."""
    from hypha_b200 import config as C

    env = {"HYPHA_SCHEDULER__JOB__ROUNDS__UPDATE_ROUNDS": "7", "HYPHA_SCHEDULER__JOB__ELASTIC__ENABLED": "false",
           "HYPHA_SCHEDULER__JOB__AGGREGATION__MODE": "collective", "HYPHA_STATUS_BRIDGE": "127.0.0.1:61000",
           "HYPHA_GATEWAY_ADDRESSES": '["tcp://10.1.2.3:4001"]', "HYPHA_DEVICE": "3", "HYPHA_NO_AUTOBUILD": "1"}
    cfg = C.builder(C.SchedulerConfig).with_env(env).build()
    job = cfg.scheduler.job
    assert job.rounds.update_rounds == 7 and job.elastic.enabled is False and job.aggregation.mode == "collective"
    assert job.rounds.avg_samples_between_updates == 1200            # siblings keep their defaults
    assert cfg.status_bridge == "127.0.0.1:61000" and cfg.gateway_addresses == ["tcp://10.1.2.3:4001"]
    w = C.builder(C.WorkerConfig).with_env(env).build()
    assert w.device == 3


def test_aggregation_weighting_is_validated_and_travels_to_the_executor(tmp_path):
    from hypha_b200 import messages as M

    p = tmp_path / "s.toml"
    p.write_text('[scheduler.job.aggregation]\nweighting = "samples"\n')
    cfg = C.builder(C.SchedulerConfig).with_toml(str(p)).build()
    assert cfg.scheduler.job.aggregation.weighting == "samples"
    C.validate(cfg)
    p.write_text('[scheduler.job.aggregation]\nweighting = "by-luck"\n')
    with pytest.raises(C.ConfigError, match="weighting"):
        C.validate(C.builder(C.SchedulerConfig).with_toml(str(p)).build())
    wire = {"model": M.Model("causal-lm", M.Fetch.uri("file:///m"), ["input_ids"]).to_wire(),
            "data": M.Fetch.scheduler("s", "d").to_wire(), "updates": M.Send.to_peers(["ps"]).to_wire(),
            "results": M.Receive.from_peers(["ps"]).to_wire(), "optimizer": M.Adam(1e-3).to_wire(), "batch_size": 4,
            "weighting": "samples"}
    t = M.TrainExecutorConfig.from_wire(wire)
    assert t.weighting == "samples" and t.to_wire()["weighting"] == "samples"
    del wire["weighting"]
    assert "weighting" not in M.TrainExecutorConfig.from_wire(wire).to_wire()      # reference job JSON unchanged


def test_job_precision_is_validated_and_travels_to_the_executor(tmp_path):
    """`[scheduler.job.precision] gemm = "mxfp8"` (BASELINE config 3 through the product path): validated, carried in
    the train executor's job JSON only when it differs from bf16, applied by the executor on GPUs only."""
    from hypha_b200 import messages as M
    from hypha_b200.executor.training import apply_precision

    p = tmp_path / "s.toml"
    p.write_text('[scheduler.job.precision]\ngemm = "mxfp8"\n')
    cfg = C.builder(C.SchedulerConfig).with_toml(str(p)).build()
    assert cfg.scheduler.job.precision.gemm == "mxfp8"
    C.validate(cfg)
    assert C.builder(C.SchedulerConfig).build().scheduler.job.precision.gemm == "bf16"
    p.write_text('[scheduler.job.precision]\ngemm = "int3"\n')
    with pytest.raises(C.ConfigError, match="precision"):
        C.validate(C.builder(C.SchedulerConfig).with_toml(str(p)).build())
    wire = {"model": M.Model("causal-lm", M.Fetch.uri("file:///m"), ["input_ids"]).to_wire(),
            "data": M.Fetch.scheduler("s", "d").to_wire(), "updates": M.Send.to_peers(["ps"]).to_wire(),
            "results": M.Receive.from_peers(["ps"]).to_wire(), "optimizer": M.Adam(1e-3).to_wire(), "batch_size": 4,
            "precision": "mxfp8"}
    t = M.TrainExecutorConfig.from_wire(wire)
    assert t.precision == "mxfp8" and t.to_wire()["precision"] == "mxfp8"
    wire["precision"] = "bf16"
    assert "precision" not in M.TrainExecutorConfig.from_wire(wire).to_wire()     # reference job JSON unchanged
    assert apply_precision("mxfp8", "cpu") == "bf16" and apply_precision(None, "cuda:0") == "bf16"
    with pytest.raises(ValueError):
        apply_precision("int3", "cpu")
