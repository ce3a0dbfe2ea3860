"""Micro-benchmark of one parameter-server round on the host: the one-pass C++ reducer
(`csrc/core/ps_reduce.hpp`) against the tensor-by-tensor torch path, GPT-2-small-sized by default.

    python bench/ps_round.py [--workers 2] [--elements 124000000] [--dir /dev/shm]

Prints one JSON line.  Files live in --dir (default: a tmpfs if there is one, so that the number is the
reducer's and not the disk's)."""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from safetensors.torch import save_file  # noqa: E402

from hypha_b200.worker import ps_round  # noqa: E402


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--workers", type=int, default=2)
    ap.add_argument("--elements", type=int, default=124_000_000)
    ap.add_argument("--tensors", type=int, default=148)
    ap.add_argument("--dir", default="/dev/shm" if os.path.isdir("/dev/shm") else None)
    ap.add_argument("--threads", type=int, default=0)
    a = ap.parse_args()
    per = a.elements // a.tensors
    out = {}
    for mode in ("native", "torch"):
        d = tempfile.mkdtemp(prefix="hypha-psbench-", dir=a.dir)
        try:
            times = []
            for rnd in range(3):                       # round 0 has no momentum / cumulative input yet
                paths = []
                g = torch.Generator().manual_seed(rnd)
                for w in range(a.workers):
                    p = os.path.join(d, f"w{w}")
                    save_file({f"t{i:03d}": torch.randn(per, generator=g) for i in range(a.tensors)}, p)
                    paths.append(p)
                t0 = time.perf_counter()
                r = ps_round.outer_round(paths, d, 0.7, 0.9, threads=a.threads, force_fallback=(mode == "torch"))
                times.append(time.perf_counter() - t0)
                assert r["native"] == (mode == "native")
            n = per * a.tensors
            moved = 4 * n * (a.workers + 2 + 3)        # reads: workers + momentum + cum; writes: u, m, cum
            out[mode] = {"ms": round(min(times[1:]) * 1e3, 1), "GBps": round(moved / min(times[1:]) / 1e9, 2),
                         "threads": r["threads"]}
        finally:
            shutil.rmtree(d, ignore_errors=True)
    print(json.dumps({"bench": "ps_round", "workers": a.workers, "elements": per * a.tensors,
                      "cpus": os.cpu_count(), **out,
                      "speedup": round(out["torch"]["ms"] / out["native"]["ms"], 2)}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
