#!/usr/bin/env python
"""k_ozaki_t<double> with digit planes 0..3 of W in tensor memory (option ozaki_tsw = 2: the TS
form of tcgen05.mma, no shared-memory read of those planes) against the default (all of W read
from shared memory, groups 4 and 5 double-buffered): parity vs NumPy on the probe shapes and
five launches of every sweep-step shape per variant (kernel durations: run under
`ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ozaki_t`, then
tools/ncu_durations.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.ozaki_t_probe import EXTRA, SMALL, SWEEP, operands, reference  # noqa: E402


def main():
    import picoquant_jl_b200  # noqa: F401
    from picoquant_jl_b200.host.b200_backend import B200Backend
    ok = True
    cases = dict(SMALL)
    cases.update(EXTRA)
    cases.update(SWEEP)
    for name, (ad, ai, bd, bi) in cases.items():
        A, B = operands(ad, bd, 7)
        ref = reference(A, ai, B, bi)
        for label, tsw in (("smem", 0), ("tmem", 2)):
            b = B200Backend(np.complex128)
            b.set_option("zgemm_ozaki", 6)
            b.set_option("fused", 0)
            b.set_option("ozaki_tsw", tsw)
            for rep in range(5 if name in SWEEP else 1):
                b.save_tensor_data("A", A)
                b.save_tensor_data("B", B)
                b.contract_tensors("A", ai, "B", bi, "C")
            got = np.asarray(b.load_tensor_data("C"))
            err = float(np.linalg.norm(got.ravel() - ref.ravel()) / np.linalg.norm(ref.ravel()))
            dbg = b.microbench("ozaki_t_debug")
            print(name, label, "rel_l2 %.3e" % err, "watchdog", dbg, flush=True)
            ok = ok and err < 1e-11 and dbg == 0
            b.close()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
