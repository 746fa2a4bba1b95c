#!/usr/bin/env python
"""BASELINE config 4: single amplitude of a Google-style RQC on a rows x cols grid
(default 6x6, depth 20, validating_RQC.jl shape), un-decomposed network, one GPU,
explicit pair plan from the deterministic harness planner (greedy, noise-free).

Reports, per dtype: device time of the whole plan as a compiled program, the per-kernel
split of an eager event-timed pass and the achieved TFLOP/s of the GEMM-shaped steps.
Parity of the same workload against the oracle is tests/test_gpu_parity.py::
test_config4_rqc_6x6_d20_amplitude (this tool does not touch oracle/)."""
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import picoquant_jl_b200  # noqa: E402,F401
from picoquant_jl_b200.host import (DSLBackend, add_input, add_output,  # noqa: E402
                                    contract_network, convert_circuit_to_network, create_RQC)
from picoquant_jl_b200.host.b200_backend import B200Backend  # noqa: E402
from picoquant_jl_b200.host.planner import greedy_plan, plan_cost  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 6
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 6
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 20
n = rows * cols
circ = create_RQC(rows, cols, depth, seed=0)


def network(backend):
    tn = convert_circuit_to_network(circ, backend, decompose=False)
    add_input(tn, "0" * n)
    add_output(tn, "0" * n)
    return tn


dsl = DSLBackend()
tn = network(dsl)
plan = greedy_plan(tn)
cost = plan_cost(tn, plan)
contract_network(tn, plan, "")
text = dsl.text()
out = {"workload": "rqc_%dx%d_d%d_seed0_amplitude" % (rows, cols, depth), "contractions": len(plan),
       "complex_macs": cost["macs"], "largest_intermediate_elems": cost["max_size"],
       "top_steps_log2_MNK": [[round(math.log2(x)) for x in s] for s in
                              sorted(cost["steps"], key=lambda s: -s[0] * s[1] * s[2])[:4]]}
print(out, flush=True)

for dtype, tag in ((np.complex128, "c128"), (np.complex64, "c64")):
    b = B200Backend(dtype)
    for key, arr in dsl.store.data.items():
        b.save_tensor_data(key, arr)
    prog = b.compile_program(text)
    best = 1e30
    for rep in range(4):
        b.sync()
        b.timer_begin()
        prog.run()
        ms = b.timer_end()
        if rep:
            best = min(best, ms)
    amp = complex(np.asarray(b.load_tensor_data("result")).reshape(-1)[0])
    r = {"program_ms": best, "tflops": 8.0 * cost["macs"] / (best * 1e-3) / 1e12,
         "arena_gib": prog.arena_bytes / 2 ** 30, "launches": prog.launches,
         "amplitude": [amp.real, amp.imag]}
    b.profile_enable(True)
    prog.run()
    prof = b.profile_read()
    b.profile_enable(False)
    r["kernels"] = {c: {"launches": v["launches"], "ms": round(v["ms"], 4),
                        "tflops": round(v["flops"] / v["ms"] / 1e9, 2) if v["flops"] else None,
                        "gbs": round(v["bytes"] / v["ms"] / 1e6, 1) if v["bytes"] else None}
                    for c, v in prof.items()}
    out[tag] = r
    print(tag, json.dumps(r), flush=True)
    prog.close()
    b.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "rqc_probe.json"), "w"), indent=1)
