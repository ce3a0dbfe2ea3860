#!/usr/bin/env python
"""Summarise an .ncu-rep (read here on the CPU box with `ncu -i`) into a markdown table."""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__cluster_dim_x", "cluster"), ("smsp__inst_executed.sum", "inst")]


def main(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [(hdr.index(k), lab, k) for k, lab in WANT if k in hdr]
    lines = [f"# ncu summary of `{path}`", "", "captured with `ncu --set full --clock-control none --import-source on` "
             "(one GPU, under gpurun); numbers under ncu are for SHARES / pipe utilisation, not for timing claims.", "",
             "| kernel | " + " | ".join(f"{lab} [{units[i]}]" if units[i] else lab for i, lab, _ in cols) + " |",
             "|---|" + "---|" * len(cols)]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        name = r[ki].split("(")[0].replace("unnamed>::", "").replace("hb::", "").replace("<unnamed>::", "")
        vals = []
        for i, lab, _ in cols:
            try:
                v = float(r[i].replace(",", ""))
                vals.append(f"{v:.4g}")
            except ValueError:
                vals.append(r[i])
        lines.append(f"| `{name[:60]}` | " + " | ".join(vals) + " |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
