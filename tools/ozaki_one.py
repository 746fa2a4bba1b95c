#!/usr/bin/env python
"""Runs the dominant sweep-step contraction of the bench workload (M=2^18, N=K=64, contracted
axes scattered over A) a few times -- the target of `ncu --set full -k regex:k_ozaki_t` (default
policy) or `-k regex:k_zgemm_skinny` (groups = -1: option ozaki_auto = 0, the DMMA kernel).
usage: ozaki_one.py c128|c64 [groups: 0 default policy, 6 / 4 forced, -1 INT8 off]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import picoquant_jl_b200  # noqa
from picoquant_jl_b200.host.b200_backend import B200Backend
dt = sys.argv[1] if len(sys.argv) > 1 else "c128"
groups = int(sys.argv[2]) if len(sys.argv) > 2 else 0
con = [3, 4, 5, 18, 20, 22]
ai, o, k = [], 0, 0
for i in range(24):
    if i in con:
        k += 1; ai.append(k)
    else:
        o += 1; ai.append(-o)
bi = list(range(6, 0, -1)) + [-(o + 1 + j) for j in range(6)]
rng = np.random.default_rng(0)
npdt = np.complex128 if dt == "c128" else np.complex64
A = (rng.standard_normal(2 ** 24) + 1j * rng.standard_normal(2 ** 24)).astype(npdt).reshape((2,) * 24, order="F")
B = (rng.standard_normal(2 ** 12) + 1j * rng.standard_normal(2 ** 12)).astype(npdt).reshape((2,) * 12, order="F")
b = B200Backend(npdt)
if groups > 0:
    b.set_option("zgemm_ozaki" if dt == "c128" else "cgemm_ozaki", groups)
elif groups < 0:
    b.set_option("ozaki_auto", 0)
for rep in range(5):
    b.save_tensor_data("A", A); b.save_tensor_data("B", B)
    b.contract_tensors("A", ai, "B", bi, "C")
b.sync()
