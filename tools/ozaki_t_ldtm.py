#!/usr/bin/env python
"""TMEM read rate (bytes per SM clock) of tcgen05.ld.32x32b.x<W> with 1, 2 or 4 warps per lane quadrant."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import picoquant_jl_b200  # noqa
from picoquant_jl_b200.host.b200_backend import B200Backend
b = B200Backend(np.complex128)
res = {}
for W in (4, 8, 16, 32):
    for warps in (4, 8, 16):
        k = "%d_%d" % (W, warps)
        res[k] = b.microbench("ozaki_t_ldtm_" + k)
        print("x%-2d %2d warps: %6.1f B/clk/SM  (%.0f clk per load per warp)" % (W, warps, res[k], warps * W * 128 / res[k]), flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "ozaki_t_ldtm.json"), "w"), indent=1)
