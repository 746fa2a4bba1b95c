#!/usr/bin/env python
"""profiles/sass_summary.txt: per-kernel counts of the SASS mnemonics that prove which hardware
units libpq_b200.so uses (cuobjdump -sass on the built library; runs without a GPU).
usage: tools/sass_summary.py [out]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "picoquant.jl_b200", "csrc", "libpq_b200.so")
out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_summary.txt")
WATCH = ["DMMA", "UTCHMMA", "UTCIMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCCP", "LDGSTS", "UTMALDG", "UTMASTG",
         "SYNCS", "LDG.E.ENL2.256", "STG.E.ENL2.256", "LDG.E.128", "STG.E.128", "DFMA", "I2F.F64.S64", "F2I", "PRMT",
         "SHFL", "BAR.SYNC", "ATOMS", "RED"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()  # noqa: E731
counts, order, cur, total = {}, [], None, collections.Counter()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        order.append(cur)
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["_instructions"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + ".") or (w.count(".") and op.startswith(w)):
                counts[cur][w] += 1
                total[w] += 1
with open(out_path, "w") as f:
    f.write("# SASS mnemonic counts per kernel of libpq_b200.so (cuobjdump -sass, sm_100a)\n")
    f.write("# DMMA = FP64 tensor pipe; UTCHMMA / UTCIMMA = tcgen05.mma kind::tf32 / kind::i8; LDTM / STTM = tcgen05.ld / st;\n")
    f.write("# LDGSTS = cp.async; UTMALDG / UTMASTG = TMA; SYNCS = mbarrier; *.ENL2.256 = 256-bit global access\n")
    f.write("# library totals: " + ", ".join("%s %d" % (w, total[w]) for w in WATCH if total[w]) + "\n\n")
    for fn in order:
        c = counts[fn]
        name = demangle(fn).replace("(anonymous namespace)::", "")
        name = re.sub(r"\(.*", "", name).replace("void ", "").replace("pq::", "")
        hits = ", ".join("%s %d" % (w, c[w]) for w in WATCH if c[w])
        f.write("%-72s %6d instr  %s\n" % (name[:72], c["_instructions"], hits))
print(open(out_path).read()[:3000])
