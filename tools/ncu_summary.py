#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into a small CSV + the top stall sites of the source page.
usage: ncu_summary.py <report.ncu-rep> <out_prefix>"""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.avg.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
with open(out + "_metrics.csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "metric", "unit", "value"])
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        for i, h in enumerate(hdr):
            if h in KEEP or (h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("not_issued") and r[i] not in ("0", "")):
                w.writerow([name, h, units[i], r[i]])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
if len(srows) > 2:
    sh = srows[1]
    isrc, isamp, iex = sh.index("Source"), sh.index("# Samples"), sh.index("Instructions Executed")
    data = [(r[isrc].strip(), int(r[isamp] or 0), int(r[iex] or 0)) for r in srows[2:] if len(r) > isamp]
    tot = sum(d[1] for d in data) or 1
    with open(out + "_hot_sass.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["sass_index", "pct_of_stall_samples", "instructions_executed", "sass"])
        for i, d in sorted(enumerate(data), key=lambda t: -t[1][1])[:60]:
            w.writerow([i, round(100.0 * d[1] / tot, 2), d[2], d[0]])
print("wrote", out + "_metrics.csv", out + "_hot_sass.csv")
