#!/usr/bin/env python
"""Probe of the INT8 tensor-core kernel (csrc/kernels_zgemm_ozaki2.cu, k_ozaki_t; forced by options
zgemm_ozaki = 6 / cgemm_ozaki = 4, default under the ozaki_auto policy).  Every stage runs in a child process
under a timeout (a wrong mbarrier phase must not hold the box):

  stage p  parity, both element types: one tile, ragged, many tiles, narrow N, short K, the
           sweep-step shapes of the bench workload (vs NumPy complex128)
  stage t  timing of the sweep-step shapes: k_ozaki_t (forced) vs the other path (ozaki_auto = 0:
           DMMA / K1 + tcgen05 3xTF32)
  stage r  phase trace of block 0 on the dominant step (PQ_OZAKI_TRACE) -> gpurun_out/ozaki_t_trace_*.bin
  stage s  one slice pair of the bench workload, amplitude vs the oracle, ms per slice

Writes gpurun_out/ozaki_t_probe.json; exit code 0 only if every parity figure is within
1e-11 (ComplexF64) / 1e-6 (ComplexF32) rel-L2 and no watchdog fired.
"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def case(rank_a, con_pos, nb_open=6):
    ai, o, k = [], 0, 0
    for i in range(rank_a):
        if i in con_pos:
            k += 1
            ai.append(k)
        else:
            o += 1
            ai.append(-o)
    nk = len(con_pos)
    bi = list(range(nk, 0, -1)) + [-(o + 1 + j) for j in range(nb_open)]
    return (2,) * rank_a, ai, (2,) * (nk + nb_open), bi


SMALL = {
    "tile_128x64x64": ((128, 64), [-1, 1], (64, 64), [-2, 1]),          # canonical A[m,k], B[n,k]
    "ragged_100x33x40": ((100, 40), [-1, 1], (33, 40), [-2, 1]),
    "two_tiles_200x64x64": ((200, 64), [-1, 1], (64, 64), [-2, 1]),
    "many_tiles_40000x17x8": ((40000, 8), [-1, 1], (17, 8), [-2, 1]),
    "k_first_64x300x24": ((64, 300), [1, -1], (64, 24), [1, -2]),       # contracted axis fastest
}
SWEEP = {
    "con_3_4_5_18_20_22": case(24, [3, 4, 5, 18, 20, 22]),
    "con_0_1_2_19_21_23": case(24, [0, 1, 2, 19, 21, 23]),
    "con_tail_18_23": case(24, [18, 19, 20, 21, 22, 23]),
    "M17_K6": case(23, [3, 4, 5, 18, 20, 22]),
    "M18_K5": case(23, [3, 4, 18, 20, 22]),
    "M18_K3_N6": case(21, [18, 19, 20]),
    "M17_N5_K6": case(23, [3, 4, 5, 18, 20, 22], nb_open=5),
}


def operands(ad, bd, seed):
    rng = np.random.default_rng(seed)
    A = (rng.standard_normal(int(np.prod(ad))) + 1j * rng.standard_normal(int(np.prod(ad))))
    B = (rng.standard_normal(int(np.prod(bd))) + 1j * rng.standard_normal(int(np.prod(bd))))
    return A.reshape(ad, order="F"), B.reshape(bd, order="F")


def reference(A, ai, B, bi):
    con = sorted(x for x in ai if x > 0)
    out = sorted((x for x in ai + bi if x < 0), reverse=True)
    letters = {}
    for x in con + out:
        letters[x] = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"[len(letters)]
    sub = "%s,%s->%s" % ("".join(letters[x] for x in ai), "".join(letters[x] for x in bi),
                         "".join(letters[x] for x in out))
    return np.einsum(sub, A, B, optimize=True)


OUT = os.path.join(ROOT, "gpurun_out", "ozaki_t_probe.json")
EXTRA = {
    "one_row_1x3x5": ((1, 5), [-1, 1], (3, 5), [-2, 1]),
    "odd_M_4097x64x64": ((4097, 64), [-1, 1], (64, 64), [-2, 1]),
    "K1_outer_5000x7": ((5000, 1), [-1, 1], (7, 1), [-2, 1]),
    "N64_K8_M2e16": case(19, [16, 17, 18]),
}
DT = {"c128": (np.complex128, "zgemm_ozaki", 6, 1e-11), "c64": (np.complex64, "cgemm_ozaki", 4, 1e-6)}


def child(stage):
    import picoquant_jl_b200  # noqa: F401
    from picoquant_jl_b200.host.b200_backend import B200Backend
    res = {}
    if stage == "p":
        cases = dict(SMALL)
        cases.update(EXTRA)
        cases.update(SWEEP)
        for name, (ad, ai, bd, bi) in cases.items():
            A, B = operands(ad, bd, 1)
            ref = reference(A, ai, B, bi)
            for dt, (npdt, opt, g, tol) in DT.items():
                b = B200Backend(npdt)
                b.set_option(opt, g)
                b.set_option("fused", 0)
                b.save_tensor_data("A", A.astype(npdt))
                b.save_tensor_data("B", B.astype(npdt))
                b.profile_enable(True)
                b.contract_tensors("A", ai, "B", bi, "C")
                prof = b.profile_read()
                got = np.asarray(b.load_tensor_data("C")).astype(np.complex128)
                err = float(np.linalg.norm(got.ravel() - ref.ravel()) / np.linalg.norm(ref.ravel()))
                dbg = b.microbench("ozaki_t_debug")
                res["%s_%s" % (name, dt)] = {"rel_l2": err, "classes": sorted(prof), "watchdog": dbg, "tol": tol}
                print(name, dt, err, sorted(prof), "watchdog", dbg, flush=True)
                b.close()
    elif stage == "t":
        for name, (ad, ai, bd, bi) in SWEEP.items():
            A, B = operands(ad, bd, 2)
            for dt, (npdt, opt, g, tol) in DT.items():
                for label, val in (("other", 0), ("int8", g)):
                    b = B200Backend(npdt)
                    b.set_option(opt, val)
                    b.set_option("ozaki_auto", 0)
                    for rep in range(5):
                        b.save_tensor_data("A", A.astype(npdt))
                        b.save_tensor_data("B", B.astype(npdt))
                        if rep == 1:
                            b.profile_enable(True)
                        b.contract_tensors("A", ai, "B", bi, "C")
                    prof = b.profile_read()
                    ms = sum(r["ms"] for r in prof.values()) / 4
                    fl = max(r["flops"] for r in prof.values()) / 4
                    by = max(r["bytes"] for r in prof.values()) / 4
                    key = "%s_%s_%s" % (name, dt, label)
                    res[key] = {"ms": ms, "tflops": fl / ms / 1e9, "gbs": by / ms / 1e6}
                    print(key, res[key], flush=True)
                    b.close()
    elif stage == "r":
        ad, ai, bd, bi = SWEEP["con_3_4_5_18_20_22"]
        A, B = operands(ad, bd, 3)
        for dt, (npdt, opt, g, tol) in DT.items():
            path = os.path.join(ROOT, "gpurun_out", "ozaki_t_trace_%s.bin" % dt)
            os.environ["PQ_OZAKI_TRACE"] = path
            b = B200Backend(npdt)
            b.set_option(opt, g)
            for rep in range(2):
                b.save_tensor_data("A", A.astype(npdt))
                b.save_tensor_data("B", B.astype(npdt))
                b.contract_tensors("A", ai, "B", bi, "C")
            b.microbench("ozaki_t_trace")
            res["trace_" + dt] = {"file": os.path.basename(path), "watchdog": b.microbench("ozaki_t_debug")}
            b.close()
    elif stage == "s":
        sys.argv = ["bench.py"]
        import bench
        from picoquant_jl_b200.host.sliced import SlicedContraction
        a = bench.parse_args()
        circ, rec, name = bench.build_workload(a)
        sample = [1, 2]
        ref = complex(np.asarray(bench.run_cpu_slices(rec, np.complex128, sample)).reshape(-1)[0])
        for dt, (npdt, opt, g, tol) in DT.items():
            for label, val in (("other", 0), ("int8", g)):
                b = B200Backend(npdt)
                b.set_option(opt, val)
                b.set_option("ozaki_auto", 0)
                sc = SlicedContraction(b, rec)
                sc.run(sample, lanes=2)
                b.sync()
                b.delete_tensor("partial_sum")
                b.timer_begin()
                sc.run(sample, lanes=2)
                ms = b.timer_end()
                got = complex(np.asarray(sc.result()).reshape(-1)[0])
                key = "slices_1_2_%s_%s" % (dt, label)
                res[key] = {"rel_err": abs(got - ref) / abs(ref), "ms_per_slice": ms / len(sample),
                            "watchdog": b.microbench("ozaki_t_debug"),
                            "tol": 1e-10 if dt == "c128" else 1e-5}
                print(key, res[key], flush=True)
                b.close()
    print("RESULT " + json.dumps(res), flush=True)


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(sys.argv[2])
        return 0
    stages = sys.argv[1:] or ["p", "t", "r", "s"]
    allres, ok = {}, True
    for st in stages:
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", st],
                                 capture_output=True, text=True, timeout=420)
        except subprocess.TimeoutExpired as e:
            print("stage %s: TIMEOUT\n%s" % (st, (e.stdout or b"")[-2000:]))
            allres[st] = {"error": "timeout"}
            ok = False
            break
        sys.stdout.write(out.stdout[-6000:])
        sys.stderr.write(out.stderr[-3000:])
        res = None
        for line in out.stdout.splitlines():
            if line.startswith("RESULT "):
                res = json.loads(line[7:])
        if out.returncode != 0 or res is None:
            allres[st] = {"error": "rc %d" % out.returncode, "stderr": out.stderr[-1500:]}
            ok = False
            break
        allres[st] = res
        for k, v in res.items():
            if v.get("watchdog"):
                ok = False
            if "rel_l2" in v and not v["rel_l2"] < v["tol"]:
                ok = False
                print("FAIL", k, v)
            if "rel_err" in v and not v["rel_err"] < v["tol"]:
                ok = False
                print("FAIL", k, v)
        if not ok and st == "p":
            break
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        json.dump(allres, f, indent=1)
    print("ozaki_t probe:", "OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
