#!/usr/bin/env python
"""SM clocks per 128x64x32 kind::i8 MMA of the k_ozaki_t schedule under different conditions
(pq_microbench "ozaki_t_rate_<mode>", see csrc/kernels_zgemm_ozaki2.cu)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import picoquant_jl_b200  # noqa
from picoquant_jl_b200.host.b200_backend import B200Backend
b = B200Backend(np.complex128)
res = {}
modes = [int(x) for x in sys.argv[1:]] or [0, 1, 2, 3, 4, 6, 8, 9, 10, 16, 24, 26]
for m in modes:
    res[str(m)] = b.microbench("ozaki_t_rate_%d" % m)
    print("mode %2d: %.1f clk per MMA" % (m, res[str(m)]), flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "ozaki_t_rate.json"), "w"), indent=1)
