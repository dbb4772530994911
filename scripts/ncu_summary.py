#!/usr/bin/env python
"""Key metrics of every kernel in an .ncu-rep (run `ncu -i X --page raw --csv > raw.csv` first, or pass the .ncu-rep)."""
import csv, subprocess, sys, io
src = sys.argv[1]
if src.endswith(".ncu-rep"):
    txt = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
else:
    txt = open(src).read()
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "smsp__cycles_active.avg", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]
STALL = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    print("-" * 100)
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k); print(f"{k:78s} {r[i]} {units[i]}")
    st = sorted(((float(r[hdr.index(h)] or 0), h) for h in STALL), reverse=True)
    for v, h in st[:8]:
        print(f"  stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):40s} {v:.3f}")
