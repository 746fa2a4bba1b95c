#!/usr/bin/env python
"""Bring-up probe for the EXPERIMENTAL int8 tensor-core ZGEMM (csrc/kernels_zgemm_ozaki.cu,
option "zgemm_ozaki" = 6 / 7).  The kernel has been compiled and its arithmetic emulated on the
host (csrc/test_lower.cpp: test_ozaki) but never run on a GPU, so this probe goes in stages and
runs every stage in a child process under a timeout -- a wrong mbarrier phase shows up as a
hang, and a hung child must not hold the box:

    gpurun --timeout 900 -- 'python tools/ozaki_probe.py'

  stage 0  pq_microbench("umma_i8_selftest"): one kind::i8 MMA on known patterns (descriptor
           encodings, TMEM mapping), then the raw INT8 issue rate at N = 32 / 64
  stage 1  one tile, canonical layouts (M=128, N=K=64), then ragged M / N / K
  stage 2  the sweep-step shapes of the bench workload (gather fused), parity vs NumPy c128
  stage 3  timing of those shapes against the DMMA kernels (option off)
  stage 4  one slice of the bench workload as a compiled program, amplitude vs the oracle
  stage 5  the ComplexF32 twin (option cgemm_ozaki = 4): parity vs the c128 reference and timing
           against the K1 + tcgen05 3xTF32 path
  stage 6  long contractions (64 < K <= 8192, k_zgemm_ozaki_kloop behind K1): parity and timing
           against the 3M DMMA kernel on the top shapes of BASELINE config 4

Writes gpurun_out/ozaki_probe.json.  Exit code 0 only if every stage that ran is within
tolerance (1e-11 rel-L2 per contraction, 1e-10 for the amplitude).
"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out", "ozaki_probe.json")


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


def child(stage):
    import picoquant_jl_b200  # noqa: F401
    from picoquant_jl_b200.host.b200_backend import B200Backend
    res = {}
    if stage == "0":
        b = B200Backend(np.complex128)
        os.environ.setdefault("PQ_OZAKI_DUMP", os.path.join(ROOT, "gpurun_out", "umma_i8_selftest.bin"))
        res["selftest_wrong_entries"] = {"wrong": b.microbench("umma_i8_selftest"),
                                         "watchdog": b.microbench("ozaki_debug")}
        print("selftest", res["selftest_wrong_entries"], flush=True)
        if res["selftest_wrong_entries"]["wrong"] == 0:
            for w in ("umma_i8_tops_n32", "umma_i8_tops_n64"):
                res[w] = {"tops": b.microbench(w)}
                print(w, res[w], flush=True)
        b.close()
    elif stage in ("1", "2"):
        cases = SMALL if stage == "1" else SWEEP
        for name, (ad, ai, bd, bi) in cases.items():
            A, B = operands(ad, bd, 1)
            ref = reference(A, ai, B, bi)
            for g in (6, 7):
                b = B200Backend(np.complex128)
                b.set_option("zgemm_ozaki", g)
                b.set_option("fused", 0)
                b.save_tensor_data("A", A)
                b.save_tensor_data("B", B)
                b.profile_enable(True)
                b.contract_tensors("A", ai, "B", bi, "C")
                prof = b.profile_read()
                got = np.asarray(b.load_tensor_data("C"))
                err = float(np.linalg.norm(got.ravel() - ref.ravel()) / np.linalg.norm(ref.ravel()))
                dbg = b.microbench("ozaki_debug")   # != 0: a barrier wait timed out (see stderr)
                res["%s_g%d" % (name, g)] = {"rel_l2": err, "classes": sorted(prof), "watchdog": dbg}
                print(name, g, err, sorted(prof), "watchdog", dbg, flush=True)
                b.close()
    elif stage == "3":
        for name, (ad, ai, bd, bi) in SWEEP.items():
            A, B = operands(ad, bd, 2)
            for label, g in (("dmma", 0), ("ozaki6", 6), ("ozaki7", 7)):
                b = B200Backend(np.complex128)
                b.set_option("zgemm_ozaki", g)
                for rep in range(4):
                    b.save_tensor_data("A", A)
                    b.save_tensor_data("B", B)
                    if rep == 1:
                        b.profile_enable(True)
                    b.contract_tensors("A", ai, "B", bi, "C")
                prof = b.profile_read()
                ms = sum(r["ms"] for r in prof.values()) / 3
                fl = max(r["flops"] for r in prof.values()) / 3
                by = max(r["bytes"] for r in prof.values()) / 3
                res["%s_%s" % (name, label)] = {"ms": ms, "tflops": fl / ms / 1e9, "gbs": by / ms / 1e6}
                print(name, label, res["%s_%s" % (name, label)], flush=True)
                b.close()
    elif stage == "4":
        import bench
        from picoquant_jl_b200.host.sliced import SlicedContraction

        class Args:
            rows, cols, depth, seed, slices = 7, 7, 24, 0, 64
        circ, rec, name = bench.build_workload(Args)
        ref = bench.run_cpu_slices(rec, np.complex128, [1, 2])
        for g in (0, 6, 7):
            b = B200Backend(np.complex128)
            b.set_option("zgemm_ozaki", g)
            sc = SlicedContraction(b, rec)
            sc.run([1, 2], "s")
            b.sync()
            b.timer_begin()
            sc.run([1, 2], "t")
            ms = b.timer_end()
            got = b.load_tensor_data("s")
            err = float(abs(got - ref) / abs(ref))
            res["slices_1_2_g%d" % g] = {"rel_err": err, "ms_per_slice": ms / 2,
                                         "watchdog": b.microbench("ozaki_debug")}
            print("slices", g, err, ms / 2, flush=True)
            b.close()
    elif stage == "5":
        for name, (ad, ai, bd, bi) in list(SMALL.items()) + list(SWEEP.items()):
            A, B = operands(ad, bd, 3)
            A, B = A.astype(np.complex64), B.astype(np.complex64)
            ref = reference(A.astype(np.complex128), ai, B.astype(np.complex128), bi)
            for label, g in (("tf32", 0), ("ozaki4", 4)):
                b = B200Backend(np.complex64)
                b.set_option("cgemm_ozaki", g)
                for rep in range(3):
                    b.save_tensor_data("A", A)
                    b.save_tensor_data("B", B)
                    if rep == 1:
                        b.profile_enable(True)
                    b.contract_tensors("A", ai, "B", bi, "C")
                prof = b.profile_read()
                ms = sum(r["ms"] for r in prof.values()) / 2
                got = np.asarray(b.load_tensor_data("C"))
                err = float(np.linalg.norm(got.ravel() - ref.ravel()) / np.linalg.norm(ref.ravel()))
                res["%s_%s" % (name, label)] = {"rel_l2_c64": err, "ms": ms, "classes": sorted(prof),
                                                "watchdog": b.microbench("ozaki_debug")}
                print(name, label, res["%s_%s" % (name, label)], flush=True)
                b.close()
    elif stage == "6":
        shapes = {"K200": ((300, 200), [-1, 1], (70, 200), [-2, 1]),
                  "cfg4_top_M13_N13_K11": ((2 ** 13, 2 ** 11), [-1, 1], (2 ** 13, 2 ** 11), [-2, 1]),
                  "cfg4_M12_N12_K8": ((2 ** 12, 2 ** 8), [-1, 1], (2 ** 12, 2 ** 8), [-2, 1])}
        for name, (ad, ai, bd, bi) in shapes.items():
            A, B = operands(ad, bd, 4)
            ref = A @ B.T
            for label, g in (("dmma", 0), ("ozaki6", 6)):
                b = B200Backend(np.complex128)
                b.set_option("zgemm_ozaki", g)
                for rep in range(3):
                    b.save_tensor_data("A", A)
                    b.save_tensor_data("B", B)
                    if rep == 1:
                        b.profile_enable(True)
                    b.contract_tensors("A", ai, "B", bi, "C")
                prof = b.profile_read()
                ms = sum(r["ms"] for r in prof.values()) / 2
                got = np.asarray(b.load_tensor_data("C"))
                err = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
                res["%s_%s" % (name, label)] = {"rel_l2": err, "ms": ms,
                                                "tflops": 8.0 * ad[0] * bd[0] * ad[1] / ms / 1e9,
                                                "watchdog": b.microbench("ozaki_debug")}
                print(name, label, res["%s_%s" % (name, label)], flush=True)
                b.close()
    print("RESULT " + json.dumps(res), flush=True)


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(sys.argv[2])
        return 0
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    summary, ok = {}, True
    for stage, limit in (("0", 120), ("1", 180), ("2", 300), ("3", 300), ("4", 420), ("5", 420), ("6", 420)):
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", stage],
                               capture_output=True, text=True, timeout=limit)
            lines = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")]
            summary["stage" + stage] = (json.loads(lines[-1][7:]) if lines else
                                        {"error": (r.stdout + r.stderr)[-2000:], "rc": r.returncode})
            if not lines:
                ok = False
        except subprocess.TimeoutExpired as e:
            summary["stage" + stage] = {"error": "timeout (hang?)",
                                        "stdout": (e.stdout or b"")[-2000:].decode(errors="replace")}
            ok = False
        json.dump(summary, open(OUT, "w"), indent=1)
        if not ok:
            break   # later stages build on the earlier ones
        for k, v in summary["stage" + stage].items():
            if "rel_l2" in v and not v["rel_l2"] < 1e-11:
                ok = False
            if "rel_err" in v and not v["rel_err"] < 1e-10:
                ok = False
            if "rel_l2_c64" in v and not v["rel_l2_c64"] < 1e-5:
                ok = False
            if "wrong" in v and v["wrong"] != 0:
                ok = False
            if v.get("watchdog"):
                ok = False
        if not ok:
            break
    print(json.dumps(summary, indent=1))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
