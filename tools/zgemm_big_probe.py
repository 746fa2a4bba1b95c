#!/usr/bin/env python
"""Times the ComplexF64 GEMM on canonical layouts (long K): 3M DMMA kernel vs the four-product
kernel, on 4096^3, the config-4 top step (M=N=2^13, K=2^11) and the reference micro-shape."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import picoquant_jl_b200  # noqa
from picoquant_jl_b200.host.b200_backend import B200Backend
rng = np.random.default_rng(0)
shapes = {
    "square_4096": ((4096, 4096), [-1, 1], (4096, 4096), [1, -2]),
    "rqc_M13_N13_K11": ((8192, 2048), [-1, 1], (2048, 8192), [1, -2]),
    "rqc_M11_N13_K9": ((2048, 512), [-1, 1], (512, 8192), [1, -2]),
}
out = {}
for mode, opts in (("3m", {"fused": 1}), ("4m", {"fused": 1, "zgemm_3m": 1})):
    b = B200Backend(np.complex128)
    for k, v in opts.items():
        b.set_option(k, v)
    for name, (ad, ai, bd, bi) in shapes.items():
        A = np.asarray(rng.standard_normal(ad) + 1j * rng.standard_normal(ad), order="F")
        B = np.asarray(rng.standard_normal(bd) + 1j * rng.standard_normal(bd), order="F")
        for rep in range(4):
            b.save_tensor_data("A", A); b.save_tensor_data("B", B)
            if rep == 1:
                b.profile_enable(True)
            b.contract_tensors("A", ai, "B", bi, "C")
        prof = b.profile_read(); b.profile_enable(False)
        rec = {c: round(r["ms"] / 3, 4) for c, r in prof.items()}
        rec["gemm_tflops"] = round(prof["gemm_tensor"]["flops"] / prof["gemm_tensor"]["ms"] / 1e9, 2)
        if name == "square_4096":
            got = b.load_tensor_data("C")
            ref = A @ B
            rec["rel_l2_vs_numpy"] = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
        out["%s_%s" % (name, mode)] = rec
        print(name, mode, rec, flush=True)
    b.close()
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "zgemm_big_probe.json"), "w"), indent=1)
