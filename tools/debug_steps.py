#!/usr/bin/env python
"""Step-by-step parity hunt: interprets one slice's .tl stream on the oracle and on
B200Backend (immediate mode) in lock-step and reports the first diverging op."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import picoquant_jl_b200  # noqa
from oracle import layer1
from picoquant_jl_b200.host import create_RQC
from picoquant_jl_b200.host.backends import parse_dsl
from picoquant_jl_b200.host.b200_backend import B200Backend
from picoquant_jl_b200.host.planner import sweep_plan
from picoquant_jl_b200.host.sliced import SlicedContraction, record_sliced_contraction

rows, cols, depth, P = [int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (7, 7, 24, 64))]
circ = create_RQC(rows, cols, depth, seed=0)
rec = record_sliced_contraction(circ, P, 1, plan_fn=lambda tn, s: sweep_plan(tn, rows, cols, sliced_bonds=s),
                                output_config="0" * circ.n_qubits)
b = B200Backend(np.complex128)
ref = {}
bad = 0
for i, (cmd, a) in enumerate(parse_dsl(rec.text)):
    if cmd == "tensor":
        d = rec.store.read(a["key"])
        ref[a["t"]] = np.asarray(d.astype(np.complex128), order="F")
        b.save_tensor_data(a["t"], d)
    elif cmd == "del":
        ref.pop(a["t"], None)
        b.delete_tensor(a["t"])
    elif cmd == "view":
        ref[a["v"]] = layer1.tensor_view(ref[a["t"]], a["axis"], a["idx"])
        b.view_tensor(a["v"], a["t"], a["axis"], a["idx"])
    elif cmd == "ncon":
        A, B = ref[a["A"]], ref[a["B"]]
        C = layer1.contract_tensors((A, B), (a["a_idx"], a["b_idx"]))
        ref[a["C"]] = C
        b.profile_enable(True)
        b.contract_tensors(a["A"], a["a_idx"], a["B"], a["b_idx"], a["C"])
        prof = b.profile_read()
        b.profile_enable(False)
        if C.size >= 1 << 10 or C.size == 1:
            got = b.load_tensor_data(a["C"])
            den = np.linalg.norm(C)
            err = np.linalg.norm(got - C) / (den if den > 0 else 1)
            if not err < 1e-10:
                bad += 1
                print("MISMATCH op %d %s: A%s %s B%s %s -> C%s kernels=%s err=%.3e |C|=%.3e"
                      % (i, a["C"], A.shape, a["a_idx"], B.shape, a["b_idx"], C.shape,
                         list(prof), err, den), flush=True)
                if bad >= 4:
                    break
            elif C.size >= 1 << 20 or C.size == 1:
                print("ok op %d numel=%d kernels=%s err=%.2e" % (i, C.size, list(prof), err), flush=True)
    elif cmd == "save":
        print("final oracle", ref[a["t"]], "device", b.load_tensor_data(a["t"]))
print("immediate-mode mismatches:", bad)
b2 = B200Backend(np.complex128)
sc = SlicedContraction(b2, rec)
for mode in (1, 0):
    b2.set_option("graph", mode)
    b2.delete_tensor("acc")
    sc.program.run(rec.view_starts(1), "acc")
    print("program mode graph_opt=%d ->" % mode, b2.load_tensor_data("acc"), "arena", sc.program.arena_bytes)
