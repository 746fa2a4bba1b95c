#!/usr/bin/env python
"""One-shot measurement script for a gpurun call: roofline denominators measured on
the box (copy GB/s, FP64 DMMA / DFMA issue peaks, cuBLAS ZGEMM / CGEMM through torch)
and per-kernel timings of the hot-path kernels at BASELINE shapes.  Writes
gpurun_out/probe.json.  Not part of the product path."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import picoquant_jl_b200  # noqa: E402,F401
from picoquant_jl_b200.host import (add_input, convert_circuit_to_network,  # noqa: E402
                                    create_qft_circuit, full_wavefunction_contraction)
from picoquant_jl_b200.host.b200_backend import B200Backend  # noqa: E402

out = {}


def timed(b, fn, reps=3):
    b.sync()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        b.sync()
        best = min(best, time.perf_counter() - t0)
    return best


def main():
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    b = B200Backend(np.complex128)
    out["copy_gbs"] = b.microbench("copy_gbs")
    out["dmma_tflops"] = b.microbench("dmma_tflops")
    out["dfma_tflops"] = b.microbench("dfma_tflops")
    try:
        import torch
        for name, dt in (("zgemm", torch.complex128), ("cgemm", torch.complex64)):
            n = 4096
            x = torch.randn(n, n, dtype=dt, device="cuda")
            y = torch.randn(n, n, dtype=dt, device="cuda")
            torch.matmul(x, y)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = 1e30
            for _ in range(4):
                e0.record()
                torch.matmul(x, y)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            out["cublas_%s_tflops" % name] = 8.0 * n ** 3 / (best * 1e-3) / 1e12
            del x, y
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001
        out["torch_error"] = repr(e)
    print(json.dumps(out), flush=True)

    # ---- permutes at 2^26 (1 GiB c128) ------------------------------------------
    n = 26
    rng = np.random.default_rng(0)
    for dtype, tag in ((np.complex128, "c128"), (np.complex64, "c64")):
        bb = B200Backend(dtype)
        state = (rng.standard_normal(2 ** n) + 0j).astype(dtype).reshape((2,) * n, order="F")
        perms = {
            "reverse": list(range(n, 0, -1)),
            "qft_final": list(range(1, n + 1, 2)) + list(range(n, 0, -2)),
            "gate_to_end": [i for i in range(1, n + 1) if i not in (8, 20)] + [8, 20],
            "rotate13": list(range(14, n + 1)) + list(range(1, 14)),
        }
        for mode, mtag in ((0, "tiled"), (1, "generic")):
            bb.set_option("permute", mode)
            for pname, perm in perms.items():
                bb.save_tensor_data("s", state)
                bb.permute_tensor("s", perm)  # warm
                bb.profile_enable(True)
                for _ in range(3):
                    bb.permute_tensor("s", perm)
                prof = bb.profile_read()
                bb.profile_enable(False)
                for cls, r in prof.items():
                    out["permute_%s_%s_%s_gbs" % (tag, pname, mtag)] = r["bytes"] / (r["ms"] * 1e-3) / 1e9
        bb.set_option("permute", 0)
        # gate application on the 2^26 state (QFT-26 hot step)
        gate = (rng.standard_normal((2, 2, 2, 2)) + 1j * rng.standard_normal((2, 2, 2, 2))).astype(dtype)
        for pos in ((1, 2), (8, 20), (25, 26)):
            a_idx, o = [], 0
            for i in range(1, n + 1):
                if i == pos[0]:
                    a_idx.append(1)
                elif i == pos[1]:
                    a_idx.append(2)
                else:
                    o += 1
                    a_idx.append(-o)
            bb.save_tensor_data("s", state)
            bb.profile_enable(True)
            for r in range(3):
                bb.save_tensor_data("g", np.asarray(gate, order="F"))
                bb.contract_tensors("s", a_idx, "g", [1, 2, -(o + 1), -(o + 2)], "s2")
                bb.save_output("s2", "s")
                bb.delete_tensor("s2")
            prof = bb.profile_read()
            bb.profile_enable(False)
            r = prof["contract_small"]
            out["gate2q_%s_pos%d_%d_gbs" % (tag, pos[0], pos[1])] = r["bytes"] / (r["ms"] * 1e-3) / 1e9
        bb.delete_tensor("s")
        # whole QFT-26 through the host mirror (python-driven, eager)
        circ = create_qft_circuit(n)
        tn = convert_circuit_to_network(circ, bb)
        add_input(tn, "0" * n)
        bb.sync()
        bb.reset_counters()
        t0 = time.perf_counter()
        full_wavefunction_contraction(tn, "vector")
        bb.sync()
        dt = time.perf_counter() - t0
        out["qft26_%s_seconds" % tag] = dt
        out["qft26_%s_launches" % tag] = bb.counters()["kernel_launches"]
        es = np.dtype(dtype).itemsize
        out["qft26_%s_alg_gbs" % tag] = (7.85e11 if es == 16 else 3.92e11) / dt / 1e9
        bb.close()
        print(json.dumps(out), flush=True)

    # ---- GEMM shapes (c128 DMMA vs SIMT; c64 SIMT) ----------------------------
    shapes = {
        "sweep_M18_N6_K6": ((2,) * 24, None, (2,) * 12, None),
        "square_4096": ((4096, 4096), [-1, 1], (4096, 4096), [1, -2]),
        "rqc_M13_N13_K11": ((8192, 2048), [-1, 1], (2048, 8192), [1, -2]),
    }
    a_idx, o, kk = [], 0, 0
    for i in range(24):
        if i % 4 == 1:
            kk += 1
            a_idx.append(kk)
        else:
            o += 1
            a_idx.append(-o)
    b_idx = [6, 5, 4, 3, 2, 1] + [-(o + 1 + j) for j in range(6)]
    shapes["sweep_M18_N6_K6"] = ((2,) * 24, a_idx, (2,) * 12, b_idx)
    for dtype, tag in ((np.complex128, "c128"), (np.complex64, "c64")):
        for gemm, gtag in ((0, "auto"), (1, "simt")):
            if tag == "c64" and gemm == 0:
                continue
            bb = B200Backend(dtype)
            bb.set_option("gemm", gemm)
            for sname, (ad, ai, bd, bi) in shapes.items():
                A = (rng.standard_normal(int(np.prod(ad))) + 0j).astype(dtype).reshape(ad, order="F")
                B = (rng.standard_normal(int(np.prod(bd))) + 0j).astype(dtype).reshape(bd, order="F")
                for rep in range(3):
                    bb.save_tensor_data("A", A)
                    bb.save_tensor_data("B", B)
                    if rep == 1:
                        bb.profile_enable(True)
                    bb.contract_tensors("A", ai, "B", bi, "C")
                prof = bb.profile_read()
                bb.profile_enable(False)
                for cls, r in prof.items():
                    key = "%s_%s_%s_%s" % (sname, tag, gtag, cls)
                    out[key + "_ms"] = r["ms"] / max(1, r["launches"])
                    if r["flops"]:
                        out[key + "_tflops"] = r["flops"] / (r["ms"] * 1e-3) / 1e12
                    else:
                        out[key + "_gbs"] = r["bytes"] / (r["ms"] * 1e-3) / 1e9
            bb.close()
            print(json.dumps(out), flush=True)

    with open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
