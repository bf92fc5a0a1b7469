#!/usr/bin/env python
"""Compact per-kernel summary of an .ncu-rep (ncu --set full): duration, DRAM bytes, pipe use,
issue rate and the top stall reasons.  Usage: tools/ncu_summary.py report.ncu-rep [regex ...]"""
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
pats = sys.argv[2:] or [
    r"^gpu__time_duration.sum$", r"^dram__bytes_(read|write).sum$", r"^launch__(registers_per_thread|grid_size|block_size|occupancy_limit.*|waves_per_multiprocessor)$",
    r"^sm__warps_active.avg.pct_of_peak_sustained_active$", r"^smsp__inst_executed.sum$", r"^smsp__issue_active.avg.pct_of_peak_sustained_active$",
    r"^sm__inst_executed_pipe_(xu|fma|alu|lsu|fp64|tmem|uniform|fmaheavy|fmalite).sum$", r"^smsp__inst_executed_pipe_(xu|fma|alu|lsu).sum$", r"^sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active$",
    r"^sm__pipe_(xu|fma|alu|fmaheavy|fmalite|shared)_cycles_active.avg.pct_of_peak_sustained_active$",
    r"^lts__t_bytes.sum$", r"^lts__t_sector_hit_rate.pct$", r"^l1tex__data_bank_conflicts_pipe_lsu.sum$", r"^smsp__average_warp.*_per_issue_active.*$",
    r"^smsp__average_warps_issue_stalled_.*_per_issue_active.ratio$", r"^sm__cycles_elapsed.max$", r"^gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed$",
    r"^l1tex__t_bytes.sum$", r"^sm__sass_inst_executed_op_shared.*sum$", r"^smsp__cycles_active.avg$"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, zip(units, r)))
    print("=" * 100)
    print(d["Kernel Name"][1][:120], "| grid", d.get("Grid Size", ("", ""))[1], "block", d.get("Block Size", ("", ""))[1])
    stalls = []
    for h in hdr:
        if any(re.search(p, h) for p in pats):
            u, v = d[h]
            if "issue_stalled" in h:
                try:
                    stalls.append((float(v), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
                continue
            print("  %-80s %-12s %s" % (h, u, v))
    for v, h in sorted(stalls, reverse=True)[:8]:
        print("  stall %-40s %.2f" % (h, v))
