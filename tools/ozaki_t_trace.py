#!/usr/bin/env python
"""Decodes the phase trace of k_ozaki_t (block 0; PQ_OZAKI_TRACE=<file>, pq_microbench
"ozaki_t_trace"): clock64 stamps per tile, printed in SM clocks relative to the first stamp.

events  P0 loads issued   P1 data + row exponent   P2 stage free (empty)   P3 planes written (full)
        M8 full seen      M9..M14 group g issued + committed
        E16 done[0] seen  E17 done[G-1] seen       E18 tile drained and stored
usage: tools/ozaki_t_trace.py gpurun_out/ozaki_t_trace_c128.bin [tiles]"""
import sys

import numpy as np

path = sys.argv[1]
ntiles = int(sys.argv[2]) if len(sys.argv) > 2 else 12
tr = np.fromfile(path, dtype=np.int64).reshape(32, 32)
t0 = tr[tr > 0].min()
names = {0: "P0", 1: "P1", 2: "P2", 3: "P3", 8: "M8", 9: "M9", 10: "M10", 11: "M11", 12: "M12", 13: "M13",
         14: "M14", 16: "E16", 17: "E17", 18: "E18"}
print("tile " + " ".join("%7s" % names[e] for e in sorted(names)))
for t in range(ntiles):
    print("%4d " % t + " ".join("%7d" % (tr[t, e] - t0) if tr[t, e] > 0 else "      -" for e in sorted(names)))
print()
for label, e in (("producer period (P3 -> P3)", 3), ("MMA period (M8 -> M8)", 8), ("epilogue period (E18 -> E18)", 18)):
    d = np.diff(tr[2:ntiles, e])
    if len(d):
        print("%-32s mean %.0f clk  min %d  max %d" % (label, d.mean(), d.min(), d.max()))
for label, a, b in (("P: load latency (P0 -> P1)", 0, 1), ("P: wait for the stage (P1 -> P2)", 1, 2),
                    ("P: slice + store (P2 -> P3)", 2, 3), ("M: issue of a tile (M8 -> last group)", 8, None),
                    ("E: done[0] -> done[last]", 16, 17), ("E: done[last] -> drained", 17, 18)):
    if b is None:
        last = max(e for e in range(9, 15) if tr[2, e] > 0)
        d = tr[2:ntiles, last] - tr[2:ntiles, a]
    else:
        d = tr[2:ntiles, b] - tr[2:ntiles, a]
    print("%-40s mean %.0f clk" % (label, d.mean()))

n = int((tr[:, 18] > 0).sum())
if n > 2 and tr[0, 19] > 0:
    clk = tr[n - 1, 18] - tr[0, 18]
    ns = tr[n - 1, 19] - tr[0, 19]
    print("SM clock during the kernel (clock64 / globaltimer over tiles 0..%d): %.0f MHz; block 0 lifetime %.1f us"
          % (n - 1, 1e3 * clk / ns, (tr[n - 1, 19] - tr[0, 19]) / 1e3))
