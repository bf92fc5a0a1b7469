#!/usr/bin/env python
"""Stall samples of one kernel of an ncu report, rolled up over consecutive SASS regions (buckets of N instructions)
and over stall reasons.  Usage: tools/ncu_regions.py report.ncu-rep kernel_regex [bucket] [lo hi]
With lo/hi (instruction indices) prints the SASS of that range with its samples."""
import csv, subprocess, sys, collections
rep, kre = sys.argv[1], sys.argv[2]
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 100
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", "regex:" + kre],
                     stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
st_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
ins = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] == "Address":
        break
    try:
        n = int(r[ix["Instructions Executed"]]); s = int(r[ix["# Samples"]])
    except ValueError:
        continue
    ins.append((r[ix["Address"]], r[ix["Source"]], n, s, {h[6:]: int(r[ix[h]] or 0) for h in st_cols}))
tot_s = sum(i[3] for i in ins); tot_n = sum(i[2] for i in ins)
print("instructions %d, executed %d, samples %d" % (len(ins), tot_n, tot_s))
allst = collections.Counter()
for i in ins:
    allst.update(i[4])
print("stall reasons overall:", ", ".join("%s %.1f%%" % (k, 100.0 * v / tot_s) for k, v in allst.most_common(10)))
if len(sys.argv) > 5:
    lo, hi_ = int(sys.argv[4]), int(sys.argv[5])
    for k in range(lo, min(hi_, len(ins))):
        a, src, n, s, st = ins[k]
        top = ",".join("%s:%d" % kv for kv in sorted(st.items(), key=lambda x: -x[1])[:3] if kv[1] > 0)
        print("%5d %6.2f%% %9d  %-90s %s" % (k, 100.0 * s / tot_s, n, src[:90], top))
    sys.exit(0)
for b0 in range(0, len(ins), bucket):
    blk = ins[b0:b0 + bucket]
    s = sum(i[3] for i in blk); n = sum(i[2] for i in blk)
    if s < 0.004 * tot_s:
        continue
    st = collections.Counter()
    ops = collections.Counter()
    for i in blk:
        st.update(i[4])
        t = i[1].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0] if t else "?"
        ops[op] += i[2]
    print("[%5d..%5d) samples %5.1f%%  exec %5.1f%%  | %s | %s" % (b0, b0 + bucket, 100.0 * s / tot_s, 100.0 * n / tot_n,
          " ".join("%s:%.0f%%" % (k, 100.0 * v / max(1, s)) for k, v in st.most_common(4)), " ".join("%s" % k for k, v in ops.most_common(5))))
