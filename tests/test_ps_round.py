"""`worker/ps_round.py` + `csrc/core/ps_reduce.hpp`: the one-pass native parameter-server round against the
tensor-by-tensor torch path and against the reference's numeric vector (parameter_server.rs:452-524)."""
import json
import os
import struct

import pytest
import torch
from safetensors.torch import load_file, save_file

from hypha_b200.worker import ps_round


def _files(tmp, n_workers, shapes, seed=0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    paths = []
    for w in range(n_workers):
        d = {k: torch.randn(*s, generator=g).to(dtype) if s else torch.randn((), generator=g).to(dtype) for k, s in shapes.items()}
        p = os.path.join(tmp, f"worker-{seed}-{w}")
        save_file(d, p)
        paths.append(p)
    return paths


def test_reference_vector_native(tmp_path):
    tmp = str(tmp_path)
    save_file({"w": torch.full((5,), 0.5)}, os.path.join(tmp, "a"))
    r = ps_round.outer_round([os.path.join(tmp, "a")], tmp, 0.1, 0.7)
    assert r["native"] and r["elements"] == 5
    torch.testing.assert_close(load_file(r["path"])["w"], torch.full((5,), 0.085))
    save_file({"w": torch.tensor([0.1, 0.2, 0.3, 0.4, 0.5])}, os.path.join(tmp, "b"))
    r = ps_round.outer_round([os.path.join(tmp, "b")], tmp, 0.1, 0.7)
    torch.testing.assert_close(load_file(r["path"])["w"], torch.tensor([0.0415, 0.0585, 0.0755, 0.0925, 0.1095]))
    torch.testing.assert_close(load_file(os.path.join(tmp, "cumulative_update"))["w"],
                               torch.tensor([0.1265, 0.1435, 0.1605, 0.1775, 0.1945]))
    assert not [f for f in os.listdir(tmp) if f.endswith(".new")]


@pytest.mark.parametrize("workers", [1, 2, 5])
def test_native_matches_fallback_over_rounds(tmp_path, workers):
    shapes = {"embed.weight": (257, 33), "ln.bias": (33,), "scalar": (), "big": (3, 300_001), "empty": (0, 4)}
    a, b = os.path.join(str(tmp_path), "native"), os.path.join(str(tmp_path), "torch")
    os.makedirs(a)
    os.makedirs(b)
    for rnd in range(3):
        pa = _files(a, workers, shapes, seed=rnd)
        pb = _files(b, workers, shapes, seed=rnd)
        ra = ps_round.outer_round(pa, a, 0.7, 0.9, threads=3)
        rb = ps_round.outer_round(pb, b, 0.7, 0.9, force_fallback=True)
        assert ra["native"] and not rb["native"] and ra["elements"] == rb["elements"]
        for name in ("gradient_update", "momentum", "cumulative_update"):
            x, y = load_file(os.path.join(a, name)), load_file(os.path.join(b, name))
            assert set(x) == set(y) == set(shapes)
            for k in x:
                assert x[k].shape == y[k].shape and x[k].dtype == torch.float32
                torch.testing.assert_close(x[k], y[k], rtol=1e-6, atol=1e-7)


def test_files_with_different_tensor_order_and_odd_alignment(tmp_path):
    """Worker files may order tensors differently and start their data at any byte offset."""
    tmp = str(tmp_path)
    t = {"a": torch.arange(7, dtype=torch.float32), "b": torch.arange(12, dtype=torch.float32).reshape(3, 4) * 0.5}

    def write(path, order, pad):
        header, blobs, off = {}, [], 0
        for k in order:
            raw = t[k].numpy().tobytes()
            header[k] = {"dtype": "F32", "shape": list(t[k].shape), "data_offsets": [off, off + len(raw)]}
            blobs.append(raw)
            off += len(raw)
        header["__metadata__"] = {"format": "pt"}
        blob = json.dumps(header).encode() + b" " * pad          # pad shifts the data start off 4-byte alignment
        with open(path, "wb") as f:
            f.write(struct.pack("<Q", len(blob)) + blob + b"".join(blobs))

    p1, p2 = os.path.join(tmp, "w1"), os.path.join(tmp, "w2")
    write(p1, ["a", "b"], 1)
    write(p2, ["b", "a"], 3)
    assert (ps_round.read_header(p1)[1] % 4, ps_round.read_header(p2)[1] % 4) != (0, 0)
    r = ps_round.outer_round([p1, p2], tmp, 1.0, 0.0)              # lr 1, mu 0: update == mean == t
    assert r["native"]
    got = load_file(r["path"])
    torch.testing.assert_close(got["a"], t["a"])
    torch.testing.assert_close(got["b"], t["b"])


def test_non_f32_and_mismatched_files_take_the_fallback(tmp_path):
    tmp = str(tmp_path)
    shapes = {"w": (16, 16), "steps": (1,)}
    paths = _files(tmp, 2, shapes, dtype=torch.bfloat16)
    r = ps_round.outer_round(paths, tmp, 0.7, 0.9)
    assert not r["native"]
    assert load_file(r["path"])["w"].dtype == torch.float32
    # now F32 files arrive but the state on disk came from the fallback: same names and shapes -> native again
    paths = _files(tmp, 2, shapes, seed=5)
    assert ps_round.outer_round(paths, tmp, 0.7, 0.9)["native"]


def test_corrupt_header_is_rejected(tmp_path):
    p = os.path.join(str(tmp_path), "bad")
    with open(p, "wb") as f:
        f.write(struct.pack("<Q", 1 << 40) + b"{}")
    with pytest.raises(ps_round.SafetensorsFormatError):
        ps_round.read_header(p)
    hdr = json.dumps({"w": {"dtype": "F32", "shape": [4], "data_offsets": [0, 16]}}).encode()
    with open(p, "wb") as f:
        f.write(struct.pack("<Q", len(hdr)) + hdr + b"\0" * 8)     # data region shorter than the header claims
    with pytest.raises(ps_round.SafetensorsFormatError):
        ps_round.read_header(p)


def test_native_round_is_faster_than_the_torch_path(tmp_path):
    """Not a benchmark, a guard: the one-pass reducer must not be slower than the multi-pass fallback
    (2 workers x 16 M fp32 elements)."""
    a, b = os.path.join(str(tmp_path), "n"), os.path.join(str(tmp_path), "t")
    os.makedirs(a)
    os.makedirs(b)
    shapes = {"w": (4096, 4096)}
    for d in (a, b):                                   # a first round so that momentum / cum exist
        ps_round.outer_round(_files(d, 2, shapes), d, 0.7, 0.9, force_fallback=(d == b))
    ta = ps_round.outer_round(_files(a, 2, shapes, seed=1), a, 0.7, 0.9)["seconds"]
    tb = ps_round.outer_round(_files(b, 2, shapes, seed=1), b, 0.7, 0.9, force_fallback=True)["seconds"]
    print(f"native {ta * 1e3:.1f} ms vs torch {tb * 1e3:.1f} ms")
    assert ta < tb * 1.5


def test_native_equals_fallback_property(tmp_path):
    """Property: for any number of workers, tensor shapes (incl. empty and scalar ones), learning rate and
    momentum, the one-pass reducer and the torch path agree on update, momentum and cumulative update over
    two consecutive rounds."""
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st

    counter = iter(range(10 ** 6))
    shape = st.lists(st.integers(0, 9), min_size=0, max_size=3).map(tuple)

    @settings(max_examples=20, deadline=None)
    @given(workers=st.integers(1, 4), shapes=st.lists(shape, min_size=1, max_size=4),
           lr=st.floats(0.01, 2.0), mu=st.floats(0.0, 0.99))
    def check(workers, shapes, lr, mu):
        i = next(counter)
        a, b = os.path.join(str(tmp_path), f"n{i}"), os.path.join(str(tmp_path), f"t{i}")
        os.makedirs(a)
        os.makedirs(b)
        named = {f"t{j}": s for j, s in enumerate(shapes)}
        for rnd in range(2):
            ra = ps_round.outer_round(_files(a, workers, named, seed=rnd), a, lr, mu, threads=2)
            rb = ps_round.outer_round(_files(b, workers, named, seed=rnd), b, lr, mu, force_fallback=True)
            assert ra["native"] and not rb["native"]
            for name in ("gradient_update", "momentum", "cumulative_update"):
                x, y = load_file(os.path.join(a, name)), load_file(os.path.join(b, name))
                assert set(x) == set(y)
                for k in x:
                    assert x[k].shape == y[k].shape
                    torch.testing.assert_close(x[k], y[k], rtol=2e-6, atol=1e-6)

    check()


def test_sample_weighted_mean(tmp_path):
    """`aggregation.weighting = "samples"`: worker files carry `__samples__`; the round takes the mean over
    samples (native and fallback agree, the helper tensor does not appear in the outputs); a file without the
    count switches the whole round back to the mean over workers."""
    a, b = os.path.join(str(tmp_path), "n"), os.path.join(str(tmp_path), "t")
    os.makedirs(a)
    os.makedirs(b)
    g = [torch.full((3, 5), 1.0), torch.full((3, 5), 4.0), torch.full((3, 5), 10.0)]
    counts = [10.0, 30.0, 60.0]

    def files(d, with_counts=(True, True, True)):
        out = []
        for i, (t, c, w) in enumerate(zip(g, counts, with_counts)):
            tensors = {"w": t.clone(), "b": t[0].clone()}
            if w:
                tensors["__samples__"] = torch.tensor([c])
            p = os.path.join(d, f"w{i}")
            save_file(tensors, p)
            out.append(p)
        return out

    ra = ps_round.outer_round(files(a), a, 1.0, 0.0)                      # lr 1, mu 0: update == weighted mean
    rb = ps_round.outer_round(files(b), b, 1.0, 0.0, force_fallback=True)
    assert ra["native"] and not rb["native"]
    assert ra["weights"] == rb["weights"] == pytest.approx([0.1, 0.3, 0.6])
    expect = 0.1 * 1.0 + 0.3 * 4.0 + 0.6 * 10.0
    for d in (a, b):
        upd = load_file(os.path.join(d, "gradient_update"))
        assert set(upd) == {"w", "b"}                                     # the helper tensor is not part of the model
        torch.testing.assert_close(upd["w"], torch.full((3, 5), expect))
        assert set(load_file(os.path.join(d, "momentum"))) == {"w", "b"}
    # one worker does not report its count -> plain mean over workers
    c = os.path.join(str(tmp_path), "u")
    os.makedirs(c)
    rc = ps_round.outer_round(files(c, (True, False, True)), c, 1.0, 0.0)
    assert rc["native"] and rc["weights"] is None
    torch.testing.assert_close(load_file(rc["path"])["w"], torch.full((3, 5), 5.0))


def test_read_header_rejects_hostile_entries(tmp_path):
    """Worker files are untrusted: every malformed header ends as SafetensorsFormatError (never KeyError / TypeError /
    a mapped region outside the file), and a non-finite sample count is ignored instead of poisoning the weights."""
    import json
    import struct

    from hypha_b200.worker import ps_round as P

    def write(header, data=b"\0" * 64):
        blob = json.dumps(header).encode()
        p = tmp_path / f"f{abs(hash(blob)) % 10**8}.safetensors"
        p.write_bytes(struct.pack("<Q", len(blob)) + blob + data)
        return str(p)

    ok = {"w": {"dtype": "F32", "shape": [4, 4], "data_offsets": [0, 64]}}
    assert list(P.read_header(write(ok))[0]) == ["w"]
    bad = [
        {"w": {"dtype": "F32", "shape": [4, 4]}},                                        # no offsets
        {"w": {"shape": [4, 4], "data_offsets": [0, 64]}},                               # no dtype
        {"w": {"dtype": "F32", "shape": [-4, -4], "data_offsets": [0, 64]}},             # negative dims (product 16)
        {"w": {"dtype": "F32", "shape": [4, 4], "data_offsets": [-64, 0]}},              # negative offset
        {"w": {"dtype": "F32", "shape": [4, 4], "data_offsets": [0, 64.0]}},             # float offset
        {"w": {"dtype": "F32", "shape": "16", "data_offsets": [0, 64]}},                 # shape not a list
        {"w": {"dtype": "F31", "shape": [16], "data_offsets": [0, 64]}},                 # unknown dtype
        {"w": {"dtype": "F32", "shape": [16], "data_offsets": [0, 64, 128]}},
        {"w": {"dtype": "F32", "shape": [8], "data_offsets": [0, 32]}, "v": {"dtype": "F32", "shape": [8], "data_offsets": [16, 48]}},
        {"w": [1, 2, 3]},
        [1, 2],
        {"w": {"dtype": "F32", "shape": [1 << 40], "data_offsets": [0, 1 << 42]}},       # beyond the file
    ]
    for h in bad:
        with pytest.raises(P.SafetensorsFormatError):
            P.read_header(write(h))
    for v in (float("inf"), float("nan")):
        h = {"__samples__": {"dtype": "F32", "shape": [1], "data_offsets": [0, 4]}}
        path = write(h, struct.pack("<f", v) + b"\0" * 60)
        head, start = P.read_header(path)
        assert P.read_samples(path, head, start) is None
