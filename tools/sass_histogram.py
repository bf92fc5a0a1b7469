#!/usr/bin/env python
"""Per-kernel opcode histogram of libgvn.so (cuobjdump -sass): the SASS evidence that the hot path is written for the
Blackwell units -- tcgen05.mma (UTCHMMA), tensor memory (LDTM / STTM), TMA (UTMALDG / UBLKCP), mbarrier traffic
(SYNCS), packed fp32 (FFMA2 / FMUL2 / FADD2) and the special-function pipe (MUFU.*).  No GPU needed.

    python tools/sass_histogram.py [libgvn.so] > profiles/r02_sass.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "guided-vae-nmf_b200", "libgvn.so")
COLS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "SYNCS", "MUFU.EX2", "MUFU.LG2", "MUFU.RCP", "MUFU.TANH",
        "FFMA2", "FMUL2", "FADD2", "FFMA", "HMMA", "LDGSTS", "LDG", "STG", "LDS", "STS", "BAR"]
sass = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], stdout=subprocess.PIPE, text=True).stdout.strip()
kern, hist = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        hist[kern]["_total"] += 1
        base = op.split(".")[0]
        for c in COLS:
            if op.startswith(c + ".") or op == c or (("." not in c) and base == c):
                hist[kern][c] += 1
sha = subprocess.run(["sha256sum", lib], stdout=subprocess.PIPE, text=True).stdout.split()[0]
print("# SASS opcode histogram of `guided-vae-nmf_b200/libgvn.so`\n")
print("`cuobjdump -sass` of the build with sha256 `%s` (regenerate: `python tools/sass_histogram.py > profiles/r02_sass.md`)." % sha)
print("Static instruction counts per kernel; only kernels with more than 200 instructions or any tensor-core / TMA opcode are listed.\n")
print("| kernel | instr | " + " | ".join(COLS) + " |")
print("|---|---:|" + "---:|" * len(COLS))
for k, h in hist.items():
    if h["_total"] < 200 and not (h["UTCHMMA"] or h["UTMALDG"] or h["UBLKCP"]):
        continue
    name = demangle(k)
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void gvn::", "", name)
    name = re.sub(r"\(.*$", "", name)
    print("| `%s` | %d | " % (name, h["_total"]) + " | ".join(str(h[c]) if h[c] else "" for c in COLS) + " |")
tot = collections.Counter()
for h in hist.values():
    tot.update(h)
print("\nWhole library: " + ", ".join("%s %d" % (c, tot[c]) for c in COLS if tot[c]) + ".")
print("`HMMA` (legacy mma.sync tensor path) and `HGMMA` (Hopper wgmma) do not occur: the dense contraction is tcgen05 only.")
