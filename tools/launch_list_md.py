#!/usr/bin/env python
"""Markdown table of an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file x.csv`): launches, time per launch and
share per kernel.  Usage: tools/launch_list_md.py launches.csv > launches.md"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ix = {k: hdr.index(k) for k in ("Kernel Name", "Block Size", "Grid Size", "Metric Name", "Metric Unit", "Metric Value")}
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    ns = float(r[ix["Metric Value"]].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6}[r[ix["Metric Unit"]]]
    name = re.sub(r"gvn::<unnamed>::|void |\(.*$", "", r[ix["Kernel Name"]])[:64]
    a = agg.setdefault(name, [0, 0.0, r[ix["Grid Size"]], r[ix["Block Size"]]])
    a[0] += 1
    a[1] += ns
total = sum(a[1] for a in agg.values())
print("%d launches, %.1f us in total (cold-cache, serialised: compare SHARES, not absolute times)\n" % (sum(a[0] for a in agg.values()), total / 1e3))
print("| kernel | launches | us per launch | share | grid | block |\n|---|---:|---:|---:|---|---|")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.1f | %.1f %% | %s | %s |" % (name, a[0], a[1] / a[0] / 1e3, 100 * a[1] / total, a[2], a[3]))
