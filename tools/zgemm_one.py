#!/usr/bin/env python
"""Runs the dominant sweep-step contraction of the bench workload (M=2^18, N=K=64, contracted
axes scattered over A) a few times -- the target of `ncu --set full -k regex:k_zgemm_fused`."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import picoquant_jl_b200  # noqa
from picoquant_jl_b200.host.b200_backend import B200Backend
con = [3, 4, 5, 18, 20, 22]
nb_open = int(sys.argv[1]) if len(sys.argv) > 1 else 6
ai, o, k = [], 0, 0
for i in range(24):
    if i in con:
        k += 1; ai.append(k)
    else:
        o += 1; ai.append(-o)
bi = list(range(6, 0, -1)) + [-(o + 1 + j) for j in range(nb_open)]
rng = np.random.default_rng(0)
A = (rng.standard_normal(2 ** 24) + 0j).reshape((2,) * 24, order="F")
B = (rng.standard_normal(2 ** (6 + nb_open)) + 0j).reshape((2,) * (6 + nb_open), order="F")
b = B200Backend(np.complex128)
for rep in range(5):
    b.save_tensor_data("A", A); b.save_tensor_data("B", B)
    b.contract_tensors("A", ai, "B", bi, "C")
b.sync()
