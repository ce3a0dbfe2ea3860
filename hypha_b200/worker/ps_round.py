"""One parameter-server round over safetensors files: mean of the workers' pseudo-gradients, momentum,
Nesterov update, cumulative update -- all outer state lives in files of the job's work dir.

Reference: crates/worker/src/executor/parameter_server.rs -- `apply_tensor_op` :331-384 (streams tensor
pairs from mmapped files), `average_op` :194-197, `update_momentum` :386-416, `nesterov` :418-446.

Fast path: every tensor of every worker file is F32 (what `extract_gradients` produces) -> the C++ core
maps all files and does the whole round in one multi-threaded pass (`csrc/core/ps_reduce.hpp`); this module
only parses / writes the JSON headers.  Anything else (integer buffers, bf16 deltas, differing key sets)
takes the tensor-by-tensor torch fallback, which keeps one tensor per worker in RAM at a time.
"""
from __future__ import annotations

import json
import os
import struct
from typing import Dict, List, Optional, Tuple

from ..ops.native import load_core

_DT_BYTES = {"F64": 8, "F32": 4, "F16": 2, "BF16": 2, "I64": 8, "I32": 4, "I16": 2, "I8": 1, "U8": 1, "BOOL": 1,
             "U16": 2, "U32": 4, "U64": 8, "F8_E4M3": 1, "F8_E5M2": 1}
MAX_HEADER = 100 << 20


class SafetensorsFormatError(ValueError):
    pass


def read_header(path: str) -> Tuple[Dict[str, dict], int]:
    """``({name: {"dtype", "shape", "data_offsets"}}, data_start)`` -- ``__metadata__`` dropped, entries in
    file order of their data."""
    size = os.path.getsize(path)
    with open(path, "rb") as f:
        raw = f.read(8)
        if len(raw) != 8:
            raise SafetensorsFormatError(f"{path}: shorter than a safetensors header")
        (n,) = struct.unpack("<Q", raw)
        if n > MAX_HEADER or 8 + n > size:
            raise SafetensorsFormatError(f"{path}: header length {n} is not plausible")
        try:
            header = json.loads(f.read(n))
        except ValueError as exc:
            raise SafetensorsFormatError(f"{path}: header is not JSON ({exc})") from exc
    if not isinstance(header, dict):
        raise SafetensorsFormatError(f"{path}: header is not a JSON object")
    header.pop("__metadata__", None)
    start = 8 + n
    spans = []
    for name, e in header.items():
        # worker-supplied and untrusted: every field is checked before anything indexes or maps the file
        if not isinstance(e, dict) or not all(k in e for k in ("dtype", "shape", "data_offsets")):
            raise SafetensorsFormatError(f"{path}: tensor {name!r} lacks dtype / shape / data_offsets")
        off, shape, dtype = e["data_offsets"], e["shape"], e["dtype"]
        if (not isinstance(off, list) or len(off) != 2 or not all(isinstance(x, int) and not isinstance(x, bool) for x in off)
                or not isinstance(shape, list) or not all(isinstance(d, int) and not isinstance(d, bool) and d >= 0 for d in shape)
                or dtype not in _DT_BYTES):
            raise SafetensorsFormatError(f"{path}: tensor {name!r} has a malformed entry")
        a, b = off
        numel = 1
        for d in shape:
            numel *= d
        if a < 0 or b < a or start + b > size or (b - a) != numel * _DT_BYTES[dtype]:
            raise SafetensorsFormatError(f"{path}: tensor {name!r} has inconsistent offsets")
        spans.append((a, b, name))
    spans.sort()
    for (a0, b0, n0), (a1, b1, n1) in zip(spans, spans[1:]):
        if a1 < b0:
            raise SafetensorsFormatError(f"{path}: tensors {n0!r} and {n1!r} overlap")
    return dict(sorted(header.items(), key=lambda kv: kv[1]["data_offsets"][0])), start


def create_f32_file(path: str, entries: List[Tuple[str, List[int]]]) -> Dict[str, int]:
    """Write the header of an all-F32 safetensors file with ``entries`` = [(name, shape)], size the file
    (sparse) and return ``{name: absolute byte offset}``."""
    header, off = {}, 0
    for name, shape in entries:
        numel = 1
        for d in shape:
            numel *= int(d)
        header[name] = {"dtype": "F32", "shape": [int(d) for d in shape], "data_offsets": [off, off + 4 * numel]}
        off += 4 * numel
    blob = json.dumps(header, separators=(",", ":")).encode()
    blob += b" " * ((8 - len(blob) % 8) % 8)          # data starts 8-byte aligned, like the stock writer
    start = 8 + len(blob)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(blob)))
        f.write(blob)
        f.truncate(start + off)
    os.chmod(path, 0o600)
    return {name: start + header[name]["data_offsets"][0] for name, _ in entries}


def _abs_offsets(header: Dict[str, dict], start: int, names: List[str]) -> List[int]:
    return [start + header[n]["data_offsets"][0] for n in names]


SAMPLES_KEY = "__samples__"       # optional F32 [1] tensor in a worker file: samples it processed this round


def read_samples(path: str, header: Dict[str, dict], start: int) -> Optional[float]:
    e = header.get(SAMPLES_KEY)
    if e is None or e["dtype"] != "F32" or (e["data_offsets"][1] - e["data_offsets"][0]) != 4:
        return None
    with open(path, "rb") as f:
        f.seek(start + e["data_offsets"][0])
        (v,) = struct.unpack("<f", f.read(4))
    import math

    return float(v) if math.isfinite(v) else None      # inf / nan would turn every weight into nan


def sample_weights(paths: List[str], heads=None) -> Optional[List[float]]:
    """``w_k = samples_k / sum`` when EVERY worker file carries a positive sample count, else None (= true
    mean).  With performance-aware pacing workers process different numbers of samples per round; weighting
    the pseudo-gradients by them is the mean over samples instead of over workers."""
    heads = heads or [read_header(p) for p in paths]
    counts = [read_samples(p, h, s) for p, (h, s) in zip(paths, heads)]
    if any(c is None or not (c > 0) for c in counts):
        return None
    total = sum(counts)
    return [c / total for c in counts]


def _native_plan(paths: List[str], momentum_path: Optional[str], cum_path: Optional[str]):
    """Headers of all files if the fast path applies, else None."""
    heads = [read_header(p) for p in paths]
    for h, _ in heads:
        if SAMPLES_KEY in h and h[SAMPLES_KEY]["dtype"] != "F32":
            return None
    heads = [({k: v for k, v in h.items() if k != SAMPLES_KEY}, s) for h, s in heads]
    first = heads[0][0]
    names = list(first)
    for h, _ in heads:
        if set(h) != set(first):
            return None
        for n in names:
            if h[n]["dtype"] != "F32" or h[n]["shape"] != first[n]["shape"]:
                return None
    state = []
    for p in (momentum_path, cum_path):
        if p is None:
            state.append(None)
            continue
        h, s = read_header(p)
        if set(h) != set(first) or any(h[n]["dtype"] != "F32" or h[n]["shape"] != first[n]["shape"] for n in names):
            return None
        state.append((h, s))
    return heads, names, state


def outer_round(paths: List[str], work_dir: str, lr: float, mu: float, *, threads: int = 0,
                force_fallback: bool = False) -> dict:
    """Run one round.  State files in ``work_dir``: ``momentum`` and ``cumulative_update`` (absent before the
    first round) are replaced atomically; the update for the workers is written to ``gradient_update``.
    Returns ``{"path": <update file>, "native": bool, "elements": n, "seconds": s}``."""
    m_path = os.path.join(work_dir, "momentum")
    c_path = os.path.join(work_dir, "cumulative_update")
    out = os.path.join(work_dir, "gradient_update")
    have_m = m_path if os.path.exists(m_path) else None
    have_c = c_path if os.path.exists(c_path) else None
    weights = sample_weights(paths)
    plan = None if force_fallback else _native_plan(paths, have_m, have_c)
    if plan is not None:
        heads, names, (mh, ch) = plan
        entries = [(n, heads[0][0][n]["shape"]) for n in names]
        out_off = create_f32_file(out, entries)
        create_f32_file(m_path + ".new", entries)
        create_f32_file(c_path + ".new", entries)
        numels = [(heads[0][0][n]["data_offsets"][1] - heads[0][0][n]["data_offsets"][0]) // 4 for n in names]
        stats = load_core().ps_fused_round(
            paths, [_abs_offsets(h, s, names) for h, s in heads], numels,
            have_m or "", _abs_offsets(mh[0], mh[1], names) if mh else [],
            have_c or "", _abs_offsets(ch[0], ch[1], names) if ch else [],
            out, m_path + ".new", c_path + ".new", [out_off[n] for n in names], float(lr), float(mu), int(threads),
            weights or [])
        os.replace(m_path + ".new", m_path)
        os.replace(c_path + ".new", c_path)
        return {"path": out, "native": True, "elements": int(stats["elements"]), "seconds": float(stats["seconds"]),
                "threads": int(stats["threads"]), "weights": weights}
    return _fallback_round(paths, out, m_path, c_path, have_m, have_c, lr, mu, weights)


def _fallback_round(paths, out, m_path, c_path, have_m, have_c, lr, mu, weights=None) -> dict:
    import time

    from safetensors import safe_open
    from safetensors.torch import load_file, save_file

    t0 = time.perf_counter()
    handles = [safe_open(p, framework="pt") for p in paths]
    prev_m = safe_open(have_m, framework="pt") if have_m else None
    cum = load_file(have_c) if have_c else {}
    new_m, upd = {}, {}
    inv = 1.0 / len(handles)
    elements = 0
    for k in handles[0].keys():
        if k == SAMPLES_KEY:
            continue
        if weights:
            acc = handles[0].get_tensor(k).float() * weights[0]
            for h, w in zip(handles[1:], weights[1:]):
                acc += h.get_tensor(k).float() * w                       # mean over samples
        else:
            acc = handles[0].get_tensor(k).float()
            for h in handles[1:]:
                acc += h.get_tensor(k).float()
            acc *= inv                                                   # true 1/N mean
        m = acc.clone() if prev_m is None or k not in prev_m.keys() else prev_m.get_tensor(k) * mu + acc
        u = (m * mu + acc) * lr
        new_m[k], upd[k] = m.contiguous(), u.contiguous()
        cum[k] = (cum[k] + u) if k in cum else u.clone()
        elements += acc.numel()
    del handles, prev_m
    save_file(upd, out)
    save_file(new_m, m_path + ".new")
    save_file({k: v.contiguous() for k, v in cum.items()}, c_path + ".new")
    os.replace(m_path + ".new", m_path)
    os.replace(c_path + ".new", c_path)
    return {"path": out, "native": False, "elements": elements, "seconds": time.perf_counter() - t0, "threads": 1,
            "weights": weights}
