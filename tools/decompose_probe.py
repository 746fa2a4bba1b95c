#!/usr/bin/env python
"""Wall time of `decompose_tensor!` on the device (pq_decompose: permute to [left | right], one-sided
Jacobi SVD, threshold, sqrt(S) split) on the bond-matrix shapes of the MPS contraction path that
the reference's benchmarks/gpu_comparison.jl:11-41 times (gate blocks 4x4 ... and growing bonds),
beside the CPU oracle (LAPACK zgesdd through NumPy) on the box's host cores.  The call returns chi,
so it is synchronous by contract: wall time is the right clock.  Writes gpurun_out/decompose_probe.json."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import picoquant_jl_b200  # noqa
from picoquant_jl_b200.host.b200_backend import B200Backend
from oracle import layer1

rng = np.random.default_rng(3)
out = {}
for m, n, rank in [(4, 4, 0), (16, 16, 0), (32, 32, 0), (64, 64, 0), (128, 128, 0), (256, 256, 0), (512, 384, 0),
                   (512, 384, 100), (1024, 1024, 128)]:
    # a matrix with a decaying spectrum, like a compressed bond (condition ~1e8)
    k = min(m, n)
    u, _ = np.linalg.qr(rng.standard_normal((m, k)) + 1j * rng.standard_normal((m, k)))
    v, _ = np.linalg.qr(rng.standard_normal((n, k)) + 1j * rng.standard_normal((n, k)))
    s = np.logspace(0, -8, k)
    a = np.asfortranarray((u * s) @ v.conj().T)
    b = B200Backend(np.complex128)
    ts = []
    for rep in range(4):
        b.save_tensor_data("T", a)
        b.sync()
        t0 = time.perf_counter()
        chi = b.decompose_tensor("T", [1], [2], threshold=1e-13, max_rank=rank, left_label="L", right_label="R")
        ts.append(time.perf_counter() - t0)
    L = np.asarray(b.load_tensor_data("L"))
    R = np.asarray(b.load_tensor_data("R"))
    t0 = time.perf_counter()
    Lc, Rc, chic = layer1.decompose_tensor(a, [1], [2], threshold=1e-13, max_rank=rank)
    tc = time.perf_counter() - t0
    err = np.linalg.norm(L @ R - Lc @ Rc) / np.linalg.norm(Lc @ Rc)
    key = "%dx%d_rank%d" % (m, n, rank)
    out[key] = {"gpu_ms": 1e3 * min(ts[1:]), "cpu_ms": 1e3 * tc, "chi": chi, "chi_cpu": int(chic),
                "rel_err_truncated_product": float(err)}
    print(key, out[key], flush=True)
    b.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "decompose_probe.json"), "w"), indent=1)
