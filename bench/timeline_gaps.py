#!/usr/bin/env python
"""Where is the GPU idle inside one Llama-3-8B step?  CUPTI kernel timeline (torch.profiler) -> union of the busy
intervals over all streams, the idle gaps between them, and which kernels sit on either side of the largest gaps.
-> gpurun_out/timeline_gaps.json"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypha_b200.models.llama import get_config  # noqa: E402
from hypha_b200.parallel.diloco import AdamWConfig, DiLoCoConfig  # noqa: E402
from hypha_b200.trainer import DiLoCoTrainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama3-8b")
    ap.add_argument("--seq-len", type=int, default=4096)
    ap.add_argument("--micro-batch", type=int, default=2)
    a = ap.parse_args()
    cfg = get_config(a.model)
    tr = DiLoCoTrainer.create(cfg, adamw=AdamWConfig(), diloco=DiLoCoConfig(inner_steps=1000), device=0)
    ids = torch.randint(0, cfg.vocab_size, (a.micro_batch, a.seq_len)).pin_memory()
    for _ in range(3):
        tr.step(ids)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        tr.step(ids)
        tr.step(ids)
        torch.cuda.synchronize()
    evs = []
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start:
            evs.append((e.time_range.start, e.time_range.end, e.name[:70]))
    evs.sort()
    # second step only: from the middle of the capture on
    t_mid = (evs[0][0] + evs[-1][1]) / 2
    first_of_second = min((s for s, _, n in evs if s > t_mid - 0.15 * (evs[-1][1] - evs[0][0]) and "embedding" in n.lower()), default=t_mid)
    evs = [x for x in evs if x[0] >= first_of_second]
    t0, t1 = evs[0][0], max(e for _, e, _ in evs)
    busy, gaps = 0.0, []
    cur_s, cur_e, last_name = evs[0][0], evs[0][1], evs[0][2]
    for s, e, n in evs[1:]:
        if s > cur_e:
            busy += cur_e - cur_s
            gaps.append((s - cur_e, last_name, n))
            cur_s, cur_e = s, e
            last_name = n
        else:
            if e > cur_e:
                cur_e, last_name = e, n
    busy += cur_e - cur_s
    span = t1 - t0
    ksum = sum(e - s for s, e, _ in evs)
    gaps.sort(reverse=True)
    by_pair = {}
    for g, a_, b_ in gaps:
        k = f"{a_[:40]} -> {b_[:40]}"
        by_pair.setdefault(k, [0.0, 0])
        by_pair[k][0] += g
        by_pair[k][1] += 1
    out = {"step_span_ms": span / 1e3, "gpu_busy_union_ms": busy / 1e3, "idle_ms": (span - busy) / 1e3,
           "idle_pct": 100 * (span - busy) / span, "kernel_time_sum_ms": ksum / 1e3,
           "overlap_ms (sum - union)": (ksum - busy) / 1e3, "num_kernels": len(evs), "num_gaps": len(gaps),
           "gaps_over_20us": sum(1 for g in gaps if g[0] > 20), "idle_in_gaps_over_20us_ms": sum(g[0] for g in gaps if g[0] > 20) / 1e3,
           "largest_gaps_us": [(round(g, 1), a_, b_) for g, a_, b_ in gaps[:25]],
           "idle_by_neighbours_ms": sorted(((round(v[0] / 1e3, 3), v[1], k) for k, v in by_pair.items()), reverse=True)[:25]}
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/timeline_gaps.json", "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
