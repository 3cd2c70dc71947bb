#!/usr/bin/env python
"""Extracts the metrics the profiles/ summaries quote from an `ncu --page raw --csv` dump.
usage: ncu -i X.ncu-rep --page raw --csv | python tools/ncu_extract.py > out.json"""
import csv
import json
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
rows = list(csv.reader(sys.stdin))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
names, units = rows[hdr], rows[hdr + 1]
out = []
for r in rows[hdr + 2:]:
    if len(r) != len(names):
        continue
    d = {"kernel": r[names.index("Kernel Name")]}
    for w in WANT:
        if w in names:
            i = names.index(w)
            d[w] = {"value": r[i], "unit": units[i]}
    out.append(d)
json.dump(out, sys.stdout, indent=1)
