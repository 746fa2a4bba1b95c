#!/usr/bin/env python
"""Probe of the thin-N ComplexF64 kernel (csrc/kernels_zgemm.cu, k_zgemm_thin: N <= 16, K <= 64,
A fragments fetched straight from the un-permuted tensor) against the tiled configuration it
replaces (option zgemm_thin = 1, k_zgemm_fused_t<128, 8>): parity vs NumPy and time per launch on
the thin sweep steps of the bench workload.  The host-timed figures carry launch overhead; for
kernel durations run it under `ncu --metrics gpu__time_duration.sum -k regex:k_zgemm` and read
the launch list with tools/ncu_durations.py.  Writes gpurun_out/thin_probe.json."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.ozaki_t_probe import case, operands, reference  # noqa: E402

SHAPES = {
    "M18_N8_K64": case(24, [3, 4, 5, 18, 20, 22], nb_open=3),
    "M18_N8_K64_tail": case(24, [18, 19, 20, 21, 22, 23], nb_open=3),
    "M18_N8_K64_head": case(24, [0, 1, 2, 19, 21, 23], nb_open=3),
    "M18_N8_K32": case(23, [3, 4, 18, 20, 22], nb_open=3),
    "M15_N8_K64": case(21, [3, 4, 5, 16, 18, 20], nb_open=3),
    "M18_N16_K64": case(24, [3, 4, 5, 18, 20, 22], nb_open=4),
    "M17_N4_K16": case(21, [3, 4, 18, 20], nb_open=2),
    "ragged_1001x5x13": ((1001, 13), [-1, 1], (5, 13), [-2, 1]),
    "ragged_kfirst_37x999x11": ((37, 999), [1, -1], (37, 11), [1, -2]),
}


def main():
    import picoquant_jl_b200  # noqa: F401
    from picoquant_jl_b200.host.b200_backend import B200Backend
    res, ok = {}, True
    for name, (ad, ai, bd, bi) in SHAPES.items():
        A, B = operands(ad, bd, 5)
        ref = reference(A, ai, B, bi)
        for label, thin in (("auto", 0), ("kfirst", 2), ("rowsfirst", 3), ("tiled", 1)):
            b = B200Backend(np.complex128)
            b.set_option("zgemm_thin", thin)
            for rep in range(5):
                b.save_tensor_data("A", A)
                b.save_tensor_data("B", B)
                if rep == 1:
                    b.profile_enable(True)
                b.contract_tensors("A", ai, "B", bi, "C")
            prof = b.profile_read()
            got = np.asarray(b.load_tensor_data("C"))
            err = float(np.linalg.norm(got.ravel() - ref.ravel()) / np.linalg.norm(ref.ravel()))
            ms = sum(r["ms"] for r in prof.values()) / 4
            by = max(r["bytes"] for r in prof.values()) / 4
            res["%s_%s" % (name, label)] = {"rel_l2": err, "us": ms * 1e3, "gbs": by / ms / 1e6,
                                            "classes": sorted(prof)}
            print(name, label, res["%s_%s" % (name, label)], flush=True)
            ok = ok and err < 1e-13
            b.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "thin_probe.json"), "w") as f:
        json.dump(res, f, indent=1)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
