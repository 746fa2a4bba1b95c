#!/usr/bin/env python
"""Times the ComplexF32 GEMM back ends (tcgen05 3xTF32 vs SIMT) and checks their error."""
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
    "skinny_M18_N6_K6": ((1 << 18, 64), [-1, 1], (64, 64), [1, -2]),
    "ref_micro_8192x8192x4096": ((4096, 2, 4096), [-1, -2, 1], (4096, 2, 4096), [1, -3, -4]),
}
out = {}
for mode, g in (("tcgen05", 0), ("simt", 1)):
    b = B200Backend(np.complex64)
    b.set_option("gemm", g)
    for name, (ad, ai, bd, bi) in shapes.items():
        A = (rng.standard_normal(ad) + 1j * rng.standard_normal(ad)).astype(np.complex64)
        B = (rng.standard_normal(bd) + 1j * rng.standard_normal(bd)).astype(np.complex64)
        A = np.asarray(A, order="F"); B = np.asarray(B, order="F")
        for rep in range(4):
            b.save_tensor_data("A", A); b.save_tensor_data("B", B)
            if rep == 1:
                b.profile_enable(True)
            b.contract_tensors("A", ai, "B", bi, "C")
        prof = b.profile_read(); b.profile_enable(False)
        rec = {c: round(r["ms"] / 3, 4) for c, r in prof.items()}
        gk = [c for c in prof if c.startswith("gemm")][0]
        rec["gemm_tflops"] = round(prof[gk]["flops"] / prof[gk]["ms"] / 1e9, 1)
        if name == "skinny_M18_N6_K6" or name == "square_4096":
            got = b.load_tensor_data("C")
            M = A.reshape(-1, A.shape[-1], order="F") if len(ad) == 2 else None
            ref = (A.astype(np.complex128).reshape(ad[0], -1, order="F") @ B.astype(np.complex128))
            rec["rel_l2_vs_f64"] = float(np.linalg.norm(got.reshape(ref.shape, order="F") - ref) / np.linalg.norm(ref))
        out["%s_%s" % (name, mode)] = rec
        print(name, mode, rec, flush=True)
    b.close()
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "cgemm_probe.json"), "w"), indent=1)
