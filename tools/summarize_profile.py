#!/usr/bin/env python
"""Turn gpurun_out/ ncu artefacts into the small text summaries committed under profiles/.

    python tools/summarize_profile.py launches gpurun_out/launches_r1.csv  > profiles/r1_launches.txt
    python tools/summarize_profile.py kernel   gpurun_out/prof.ncu-rep     > profiles/r1_scan_tc.txt
"""
import csv
import subprocess
import sys
from collections import defaultdict

METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if r]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    d = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except (ValueError, IndexError):
            continue
        d[r[ki].split("(")[0]][0] += 1
        d[r[ki].split("(")[0]][1] += v
    tot = sum(v[1] for v in d.values())
    print(f"# ncu --metrics gpu__time_duration.sum launch list: {path}")
    print("# (cold-cache, serialised launches: compare SHARES, not absolute times)")
    print(f"{'kernel':40s} {'launches':>8s} {'total ms':>10s} {'share':>7s}")
    for k, v in sorted(d.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:40s} {v[0]:8d} {v[1] / 1e6:10.3f} {v[1] / tot * 100:6.1f}%")


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full summary: {path}")
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        print(f"\nkernel: {r[ki]}")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"  {m:78s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
