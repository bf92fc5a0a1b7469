#!/usr/bin/env python
"""Per-source-line roll-up of an ncu source page: executed warp instructions and stall samples
per line of the .cu file.  Usage: tools/ncu_source.py report.ncu-rep kernel_regex [top_n]"""
import csv, subprocess, sys, collections, re
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda" if False else "sass", "-k", "regex:" + kre],
                     stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# find header
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
tot_inst = tot_samp = 0
ops = collections.Counter(); samp_ops = collections.Counter()
lines = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] == "Address":
        break
    try:
        n = int(r[ix["Instructions Executed"]]); s = int(r[ix["# Samples"]])
    except ValueError:
        continue
    sass = r[ix["Source"]]
    op = sass.split()[0] if not sass.startswith("@") else sass.split()[1]
    op = op.split(".")[0]
    ops[op] += n; samp_ops[op] += s
    tot_inst += n; tot_samp += s
    st = {h[6:]: int(r[i] or 0) for h, i in ix.items() if h.startswith("stall_") and "Not Issued" not in h and r[i].isdigit()}
    top_st = ",".join("%s:%d" % kv for kv in sorted(st.items(), key=lambda x: -x[1])[:3] if kv[1] > 0)
    lines.append((s, n, r[ix["Address"]], sass + "   | " + top_st))
print("total warp instructions %d, samples %d" % (tot_inst, tot_samp))
print("--- by opcode (instructions | samples)")
for op, n in ops.most_common(25):
    print("  %-12s %6.2f%%  %6.2f%%" % (op, 100.0 * n / tot_inst, 100.0 * samp_ops[op] / max(1, tot_samp)))
print("--- hottest SASS instructions by stall samples")
for s, n, a, sass in sorted(lines, reverse=True)[:top]:
    print("  %6.2f%% samp %6.2f%% inst  %s  %s" % (100.0 * s / max(1, tot_samp), 100.0 * n / tot_inst, a[-5:], sass[:150]))
