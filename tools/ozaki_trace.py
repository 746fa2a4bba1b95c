#!/usr/bin/env python
"""Phase breakdown of k_zgemm_ozaki on the dominant sweep-step shape (block 0's clock64 stamps).

    PQ_OZAKI_TRACE=gpurun_out/oz_trace.bin python tools/ozaki_trace.py        (on the GPU box)
    python tools/ozaki_trace.py --decode gpurun_out/oz_trace.bin              (anywhere)

Events per tile (worker thread 0 unless noted): 0 tile start, 1 gather done (data in registers),
2 row exponents known, 3 planes written, 4 `planes` arrive, 5+16h+g done[g] observed,
13+16h+g group g drained (first column block), 40+h pass finished; MMA thread: 48 planes seen,
49+7h+g group g issued + committed."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def decode(path, ghz=1.965):
    t = np.fromfile(path, dtype=np.int64).reshape(16, 64)
    for tile in range(1, 6):
        r = t[tile]
        if r[0] == 0:
            continue
        t0 = r[0]
        us = lambda e: (r[e] - t0) / ghz / 1e3 if r[e] else float("nan")   # noqa: E731
        print("tile %d: gather %.2f  exp-barrier %.2f  sliced %.2f  arrive %.2f | mma sees planes %.2f"
              % (tile, us(1), us(2), us(3), us(4), us(48)))
        for h in range(2):
            print("   pass %d: issued " % h + " ".join("%.2f" % us(49 + 7 * h + g) for g in range(7) if r[49 + 7 * h + g]))
            print("           done   " + " ".join("%.2f" % us(5 + 16 * h + g) for g in range(7) if r[5 + 16 * h + g]))
            print("           drained " + " ".join("%.2f" % us(13 + 16 * h + g) for g in range(7) if r[13 + 16 * h + g]))
            print("           pass end %.2f" % us(40 + h))
        nxt = t[tile + 1][0]
        if nxt:
            print("   tile period %.2f us" % ((nxt - t0) / ghz / 1e3))


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--decode":
        decode(sys.argv[2])
        return
    import picoquant_jl_b200  # noqa: F401
    from picoquant_jl_b200.host.b200_backend import B200Backend
    from ozaki_probe import SWEEP, operands
    out = os.environ.setdefault("PQ_OZAKI_TRACE", os.path.join(ROOT, "gpurun_out", "oz_trace.bin"))
    ad, ai, bd, bi = SWEEP["con_3_4_5_18_20_22"]
    A, B = operands(ad, bd, 2)
    b = B200Backend(np.complex128)
    b.set_option("zgemm_ozaki", int(os.environ.get("OZ_G", "6")))
    for rep in range(3):
        b.save_tensor_data("A", A)
        b.save_tensor_data("B", B)
        b.contract_tensors("A", ai, "B", bi, "C")
    b.sync()
    b.microbench("ozaki_trace")
    decode(out)


if __name__ == "__main__":
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    main()
