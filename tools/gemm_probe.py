#!/usr/bin/env python
"""Times the ZGEMM variants on the sweep-step shapes of the bench workload."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import picoquant_jl_b200  # noqa
from picoquant_jl_b200.host.b200_backend import B200Backend

rng = np.random.default_rng(0)
def case(rank_a, con_pos, nb_open=6):
    ai, o, k = [], 0, 0
    for i in range(rank_a):
        if i in con_pos:
            k += 1; ai.append(k)
        else:
            o += 1; ai.append(-o)
    nk = len(con_pos)
    bi = list(range(nk, 0, -1)) + [-(o + 1 + j) for j in range(nb_open)]
    return (2,) * rank_a, ai, (2,) * (nk + nb_open), bi
cases = {
    "con_3_4_5_18_20_22": case(24, [3, 4, 5, 18, 20, 22]),
    "con_0_1_2_19_21_23": case(24, [0, 1, 2, 19, 21, 23]),
    "con_15_16_17_19_21_23": case(24, [15, 16, 17, 19, 21, 23]),
    "con_tail_18_23": case(24, [18, 19, 20, 21, 22, 23]),
    "M17_K6": case(23, [3, 4, 5, 18, 20, 22]),
    "M18_K5": case(23, [3, 4, 18, 20, 22]),
    "M18_K3_N6": case(21, [18, 19, 20]),
    "M18_N3_K6": case(24, [3, 4, 5, 18, 20, 22], nb_open=3),
    "M18_N3_K6_low": case(24, [0, 1, 2, 19, 21, 23], nb_open=3),
    "M18_N4_K5": case(23, [3, 4, 18, 20, 22], nb_open=4),
    "dot_K21": ((2,) * 21, list(range(1, 22)), (2,) * 21, [((i * 5) % 21) + 1 for i in range(21)]),
}
out = {}
for mode, opts in (("fused", {}), ("fused_rowfirst", {"zgemm_kfirst": 1}), ("ttgt", {"fused": 1})):
    b = B200Backend(np.complex128)
    for k, v in opts.items():
        b.set_option(k, v)
    for name, (ad, ai, bd, bi) in cases.items():
        A = (rng.standard_normal(2 ** len(ad)) + 0j).reshape(ad, order="F")
        B = (rng.standard_normal(2 ** len(bd)) + 0j).reshape(bd, order="F")
        for rep in range(4):
            b.save_tensor_data("A", A); b.save_tensor_data("B", B)
            if rep == 1:
                b.profile_enable(True)
            b.contract_tensors("A", ai, "B", bi, "C")
        prof = b.profile_read(); b.profile_enable(False)
        tot_ms = sum(r["ms"] for r in prof.values()) / 3
        fl = max(r["flops"] for r in prof.values()) / 3
        g = prof.get("gemm_tensor")
        by = max(r["bytes"] for r in prof.values()) / 3
        out["%s_%s" % (name, mode)] = {"total_ms": round(tot_ms, 4), "eff_tflops": round(fl / tot_ms / 1e9, 2), "eff_gbs": round(by / tot_ms / 1e6, 0),
                                       "gemm_tflops": round(g["flops"] / g["ms"] / 1e9, 2) if g else None,
                                       "kernels": {c: round(r["ms"] / 3, 4) for c, r in prof.items()}}
        print(name, mode, out["%s_%s" % (name, mode)], flush=True)
    b.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "gemm_probe.json"), "w"), indent=1)
