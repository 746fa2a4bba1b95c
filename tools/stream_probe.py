#!/usr/bin/env python
"""What the memory system gives the sweep-step access pattern with no arithmetic at all
(pq_microbench "stream_<mode>_<ctas>", csrc/kernels_misc.cu): a copy of 256 MB in 64 KB tiles with
the read side laid out like A of the dominant sweep step and / or the write side like C[m + M n]."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import picoquant_jl_b200  # noqa
from picoquant_jl_b200.host.b200_backend import B200Backend
b = B200Backend(np.complex128)
names = {7: "gathered reads in the producers' lane order (4 x 128 B per warp instruction), scattered writes", 0: "contiguous reads, contiguous writes", 1: "gathered reads (8 x 8 KB), contiguous writes",
         2: "contiguous reads, scattered writes (64 x 1 KB, 4 MB apart)", 3: "gathered reads, scattered writes (the sweep step)"}
res = {"copy_gbs": b.microbench("copy_gbs")}
print("plain copy: %.0f GB/s" % res["copy_gbs"])
for mode in (0, 1, 2, 3, 7):
    for ctas in (1, 2, 4, 8):
        k = "stream_%d_%d" % (mode, ctas)
        res[k] = b.microbench(k)
        print("%-60s %d CTAs/SM (%3d KB in flight per SM): %.0f GB/s" % (names[mode], ctas, 64 * ctas, res[k]), flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "stream_probe.json"), "w"), indent=1)
