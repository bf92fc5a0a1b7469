#!/usr/bin/env python
"""Basic-block view of an ncu source page: runs of SASS instructions with the same executed count.
Usage: tools/ncu_blocks.py report.ncu-rep kernel_regex [min_share_percent]"""
import csv, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
minp = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", "regex:" + kre],
                     stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
ins = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] == "Address":
        break
    try:
        ins.append((r[0], r[ix["Source"]], int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])))
    except ValueError:
        pass
tot = sum(i[2] for i in ins); ts = sum(i[3] for i in ins)
blocks = []
cur = None
for a, s, n, sm in ins:
    if cur is None or n != cur["n"]:
        cur = dict(a=a, n=n, k=0, samp=0, ops={})
        blocks.append(cur)
    cur["k"] += 1; cur["samp"] += sm
    op = (s.split()[1] if s.startswith("@") else s.split()[0]).split(".")[0]
    cur["ops"][op] = cur["ops"].get(op, 0) + 1
print("total %d warp-instr, %d samples" % (tot, ts))
for b in blocks:
    share = 100.0 * b["n"] * b["k"] / tot
    if share >= minp or 100.0 * b["samp"] / max(1, ts) >= minp:
        top = sorted(b["ops"].items(), key=lambda x: -x[1])[:7]
        print("%s  exec %9d x %4d instr = %5.2f%% inst, %5.2f%% samp | %s" % (b["a"][-5:], b["n"], b["k"], share, 100.0 * b["samp"] / max(1, ts),
                                                                         " ".join("%s:%d" % t for t in top)))
