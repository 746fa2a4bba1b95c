#!/usr/bin/env python
"""QFT-n full-wavefunction timing (config 3) through the host mirror, eager mode."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import picoquant_jl_b200  # noqa
from picoquant_jl_b200.host import (add_input, convert_circuit_to_network, create_qft_circuit,
                                    full_wavefunction_contraction, DSLBackend)
from picoquant_jl_b200.host.b200_backend import B200Backend
n = int(sys.argv[1]) if len(sys.argv) > 1 else 26
out = {}
for dtype, tag in ((np.complex128, "c128"), (np.complex64, "c64")):
    es = np.dtype(dtype).itemsize
    alg = (7.85e11 if es == 16 else 3.92e11) if n == 26 else None
    for rep in range(3):
        b = B200Backend(dtype)
        circ = create_qft_circuit(n)
        tn = convert_circuit_to_network(circ, b)
        add_input(tn, "0" * n)
        b.sync(); b.reset_counters()
        if rep == 2:
            b.profile_enable(True)
        t0 = time.perf_counter()
        b.timer_begin()
        full_wavefunction_contraction(tn, "vector")
        t_host = time.perf_counter() - t0
        ms = b.timer_end()
        dt = time.perf_counter() - t0
        print(tag, "rep", rep, "wall %.4f s host-issue %.4f s device %.4f s" % (dt, t_host, ms / 1e3),
              "alg GB/s %.0f" % (alg / dt / 1e9) if alg else "", flush=True)
        if rep == 2:
            prof = b.profile_read()
            for k, v in prof.items():
                print("   ", k, "n=%d ms=%.2f GB/s=%.0f" % (v["launches"], v["ms"], v["bytes"] / v["ms"] / 1e6))
        else:
            out["qft%d_%s_seconds" % (n, tag)] = dt
        b.close()
    # the same flow as one compiled program (DSL stream -> CUDA graph)
    dsl = DSLBackend()
    tn = convert_circuit_to_network(create_qft_circuit(n), dsl)
    add_input(tn, "0" * n)
    full_wavefunction_contraction(tn, "vector")
    b = B200Backend(dtype)
    for key, arr in dsl.store.data.items():
        b.save_tensor_data(key, arr)
    prog = b.compile_program(dsl.text())
    for rep in range(3):
        b.sync(); b.timer_begin(); prog.run(); ms = b.timer_end()
        print(tag, "program rep", rep, "device %.4f s" % (ms / 1e3), "alg GB/s %.0f" % (alg / ms / 1e6) if alg else "",
              "arena GiB %.2f" % (prog.arena_bytes / 2 ** 30), flush=True)
    out["qft%d_%s_program_seconds" % (n, tag)] = ms / 1e3
    b.close()
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "qft_probe.json"), "w"), indent=1)
