#!/usr/bin/env python
"""Offline diagnosis of a failed `pq_microbench("umma_i8_selftest")`.

The self-test (csrc/kernels_zgemm_ozaki.cu: k_umma_i8_selftest) writes known int8 patterns
into shared memory in the layout the Ozaki kernel uses, issues ONE 128 x 32 x 32 kind::i8
MMA and dumps the int32 accumulators (PQ_OZAKI_DUMP, row-major 128 x 32).  This script
rebuilds the shared-memory image and evaluates the product under a few hypotheses about how
the tensor core interpreted the descriptors, and says which one matches the dump:

    python tools/ozaki_selftest_decode.py gpurun_out/umma_i8_selftest.bin
"""
import sys

import numpy as np

TM, TN, TK = 128, 32, 32


def pat_a(r, k):
    return (r * 7 + k * 3 + r // 64) % 127 - 63


def pat_b(c, k):
    return (c * 5 + k * 11 + 1) % 127 - 63


def image(rows, pat):
    """bytes as the kernel stores them: off = (k/16)*rows*16 + (row/8)*128 + (row%8)*16 + k%16"""
    img = np.zeros(rows * TK, dtype=np.int8)
    for r in range(rows):
        for k in range(TK):
            img[(k // 16) * rows * 16 + (r // 8) * 128 + (r % 8) * 16 + k % 16] = pat(r, k)
    return img


def read(img, rows, lbo, sbo, unsigned=False, kmax=TK):
    """operand matrix as the tensor core would fetch it with the given LBO / SBO"""
    m = np.zeros((rows, TK), dtype=np.int64)
    for r in range(rows):
        for k in range(kmax):
            off = (k // 16) * lbo + (r // 8) * sbo + (r % 8) * 16 + k % 16
            v = int(img[off % img.size])
            m[r, k] = v & 0xFF if unsigned else v
    return m


def main():
    got = np.fromfile(sys.argv[1], dtype=np.int32).reshape(TM, TN)
    ia, ib = image(TM, pat_a), image(TN, pat_b)
    hyps = {
        "as designed (LBO = rows*16 between K chunks, SBO = 128)":
            (read(ia, TM, TM * 16, 128), read(ib, TN, TN * 16, 128)),
        "LBO and SBO swapped": (read(ia, TM, 128, TM * 16), read(ib, TN, 128, TN * 16)),
        "operands read as UNSIGNED 8-bit": (read(ia, TM, TM * 16, 128, True), read(ib, TN, TN * 16, 128, True)),
        "A signed, B unsigned": (read(ia, TM, TM * 16, 128), read(ib, TN, TN * 16, 128, True)),
        "A unsigned, B signed": (read(ia, TM, TM * 16, 128, True), read(ib, TN, TN * 16, 128)),
        "only the first 16 k (one core matrix) consumed":
            (read(ia, TM, TM * 16, 128, kmax=16), read(ib, TN, TN * 16, 128, kmax=16)),
    }
    for name, (a, b) in hyps.items():
        want = a @ b.T
        for label, w in (("", want), (" (result transposed within 32x32 blocks)", None)):
            if w is None:
                continue
            bad = int((got != w).sum())
            print("%-70s %5d wrong of %d" % (name + label, bad, got.size))
    want = hyps["as designed (LBO = rows*16 between K chunks, SBO = 128)"]
    want = want[0] @ want[1].T
    rows_ok = [(got[r] == want[r]).all() for r in range(TM)]
    print("rows matching the design:", sum(rows_ok), "of", TM,
          "(first wrong row: %s)" % (rows_ok.index(False) if False in rows_ok else None))
    # lane permutations: does every dumped row equal SOME expected row?
    index = {tuple(want[r]): r for r in range(TM)}
    perm = [index.get(tuple(got[r]), -1) for r in range(TM)]
    if perm != list(range(TM)) and all(p >= 0 for p in perm):
        print("rows are a permutation of the expected ones (TMEM lane mapping):", perm[:16], "...")
    print("all zero:", bool((got == 0).all()), " all 0xffffffff (never written):", bool((got == -1).all()))


if __name__ == "__main__":
    main()
