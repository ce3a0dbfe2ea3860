"""Wire-format tests for hypha_b200.messages (reference: crates/messages/src/lib.rs; the reference has
serde derive + a few validation tests, here every message round-trips through CBOR and JSON and the
validation rules of Fetch / Send / Receive are checked).  Property tests use hypothesis like the
reference's proptest blocks."""
import json

import pytest
from hypothesis import given, settings, strategies as st

from hypha_b200 import messages as M

peer_ids = st.text(alphabet="abcdefghijklmnopqrstuvwxyz0123456789-", min_size=1, max_size=24)
floats = st.floats(min_value=0.0, max_value=1e6, allow_nan=False, allow_infinity=False)


def roundtrip(obj, cls=None):
    cls = cls or type(obj)
    wire = obj.to_wire()
    assert M.decode_cbor(M.encode_cbor(wire)) == wire                 # CBOR-stable
    assert json.loads(M.encode_json(wire)) == json.loads(json.dumps(wire))
    back = cls.from_wire(M.decode_cbor(M.encode_cbor(wire)))
    assert back.to_wire() == wire
    return back


@settings(max_examples=60, deadline=None)
@given(gpu=floats, cpu=floats, mem=floats, sto=floats)
def test_resources_roundtrip(gpu, cpu, mem, sto):
    r = roundtrip(M.Resources(gpu=gpu, cpu=cpu, storage=sto, memory=mem))
    assert (r.gpu, r.cpu, r.storage, r.memory) == (gpu, cpu, sto, mem)


@settings(max_examples=40, deadline=None)
@given(peers=st.lists(peer_ids, min_size=1, max_size=5, unique=True), ds=peer_ids, idx=st.integers(0, 10 ** 6))
def test_reference_variants_roundtrip(peers, ds, idx):
    roundtrip(M.Fetch.uri("https://example.org/model.safetensors"), M.Fetch)
    roundtrip(M.Fetch.huggingface("org/repo", "main", ["config.json", "model.safetensors"], None), M.Fetch)
    roundtrip(M.Fetch.scheduler(peers[0], ds), M.Fetch)
    roundtrip(M.Fetch.data_peer(peers[0], M.DataSlice(ds, idx)), M.Fetch)
    s = roundtrip(M.Send.to_peers(peers), M.Send)
    assert s.peers == peers
    r = roundtrip(M.Receive.from_peers(peers), M.Receive)
    assert r.get_peers() == peers


def test_send_and_receive_validation():
    """messages/src/lib.rs:334-417: Send/Receive only from `peers`; Receive only with strategy All."""
    with pytest.raises(M.MessageError):
        M.Send.from_wire(M.Fetch.uri("http://x").to_wire())
    with pytest.raises(M.MessageError):
        M.Receive.from_wire(M.Fetch.scheduler("p", "d").to_wire())
    one = M.Send.to_peers(["a", "b"], M.SelectionStrategy.ONE).to_wire()
    with pytest.raises(M.MessageError):
        M.Receive.from_wire(one)
    with pytest.raises(M.MessageError):
        M.Reference.from_wire({"type": "carrier-pigeon"})


def _train_config(**extra):
    return M.TrainExecutorConfig(
        model=M.Model("causal-lm", M.Fetch.huggingface("org/m", None, ["config.json"], None), ["input_ids", "labels"]),
        data=M.Fetch.scheduler("sched", "corpus"), updates=M.Send.to_peers(["ps"]),
        results=M.Receive.from_peers(["ps"]), optimizer=M.Adam(1e-3, (0.9, 0.95), 1e-8), batch_size=4, **extra)


def test_job_spec_roundtrip_with_b200_extensions():
    base = _train_config()
    wire = base.to_wire()
    assert "collective" not in wire and "join" not in wire          # absent unless used: reference-compatible
    assert set(wire["optimizer"]) == {"learning-rate", "betas", "epsilon"}       # kebab-case (lib.rs:652-658)
    coll = _train_config(collective={"rank": 1, "world": 8, "rendezvous": "tcp://127.0.0.1:29400",
                                     "outer": M.Nesterov(0.7, 0.9).to_wire()})
    spec = M.JobSpec(M.new_id(), M.Executor.train(M.TRAIN_EXECUTOR_NAME, coll))
    back = roundtrip(spec)
    assert back.executor.config.collective["world"] == 8
    joiner = _train_config(join={"after_round": 3, "replaces": "worker-c"})
    assert roundtrip(joiner).join == {"after_round": 3, "replaces": "worker-c"}
    agg = M.AggregateExecutorConfig(M.Receive.from_peers(["a", "b"]), M.Send.to_peers(["a", "b"]), M.Nesterov(0.7, 0.9))
    spec2 = roundtrip(M.JobSpec(M.new_id(), M.Executor.aggregate(M.AGGREGATE_EXECUTOR_NAME, agg)))
    assert spec2.executor.descriptor.cls == "aggregate"


def test_model_and_preprocessor_types_are_validated():
    assert len(M.MODEL_TYPES) == 38                                   # the 38 HF auto-classes (lib.rs:419-460)
    with pytest.raises(M.MessageError):
        M.Model("poetry", M.Fetch.uri("http://x"), []).to_wire() and M.Model.from_wire(
            {"task": "poetry", "artifact": M.Fetch.uri("http://x").to_wire(), "input-names": []})


@settings(max_examples=60, deadline=None)
@given(kind=st.sampled_from(["update", "updated", "update-received"]))
def test_progress_string_variants(kind):
    assert roundtrip(M.Progress(kind)).kind == kind


@settings(max_examples=60, deadline=None)
@given(bs=st.integers(0, 10 ** 6), rnd=st.integers(0, 10 ** 6),
       metrics=st.dictionaries(st.text(min_size=1, max_size=12), floats, max_size=4))
def test_progress_status_and_metrics(bs, rnd, metrics):
    assert roundtrip(M.Progress("status", batch_size=bs)).batch_size == bs
    m = roundtrip(M.Progress("metrics", round=rnd, metrics=metrics))
    assert m.round == rnd and m.metrics == {k: float(v) for k, v in metrics.items()}
    # extra keys in a status body are ignored (the reference executor sends a bogus round_time: quirk Q2)
    assert M.Progress.from_wire({"status": {"batch_size": bs, "round_time": -5}}).batch_size == bs
    with pytest.raises(M.MessageError):
        M.Progress.from_wire("upgrade")


def test_api_request_variants_and_membership_extension():
    for req in (M.ApiRequest.renew_lease("L1"), M.ApiRequest.data("corpus"),
                M.ApiRequest.job_status("T1", M.JobStatus.RUNNING),
                M.ApiRequest.worker_offer("L1", "R1", 1.5, M.Resources(gpu=1), 1234.5),
                M.ApiRequest.parameter_pull("J", "k"), M.ApiRequest.parameter_push("J", "k", "sid", 10),
                M.ApiRequest.job_membership("J", ["a", "b"])):
        back = roundtrip(req)
        assert back.kind == req.kind
    with pytest.raises(M.MessageError):
        M.ApiRequest.from_wire({"SelfDestruct": {}})
    assert M.ApiResponse.renewed("L1", 12.5).variant()[0] == "Renewed"
    assert M.ApiResponse.renew_failed().variant() == ("Failed", None)
    assert M.ApiResponse.job_membership(True).body == {"known": True}


@settings(max_examples=30, deadline=None)
@given(ds=st.text(alphabet="abcdefghijklmnopqrstuvwxyz-_", min_size=1, max_size=40), idx=st.integers(0, 2 ** 31))
def test_pull_header_framing(ds, idx):
    """8-byte little-endian length + JSON (stream_pull.rs:28,73-98); oversize headers are rejected."""
    import io

    raw = M.encode_pull_header(M.DataSlice(ds, idx))
    assert int.from_bytes(raw[:8], "little") == len(raw) - 8
    back = M.decode_pull_header(io.BytesIO(raw))
    assert (back.dataset, back.index) == (ds, idx)
    with pytest.raises(M.MessageError):
        M.decode_pull_header(io.BytesIO((2 << 20).to_bytes(8, "little") + b"{}"))
