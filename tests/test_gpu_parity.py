"""GPU parity tests: the CUDA path behind the C ABI (B200Backend) against the CPU
oracle on the same seeded inputs, plus the reference's own known-answer tests and
size-independent properties at BASELINE sizes.  Tolerances are the north-star's:
rel-L2 <= 1e-10 for ComplexF64 and <= 1e-5 for ComplexF32 (checked against the
F64 oracle downcast); pure data movement (permute / view) must be bit-exact."""
import random

import numpy as np
import pytest

from helpers import TOL, golden_qasm, load_golden_json, rel_l2, statevector, switch_endianness
from oracle import layer1
from oracle.interactive import OracleBackend, execute_dsl
from picoquant_jl_b200.host import (Circuit, DSLBackend, TensorNetworkCircuit, add_gate,
                                    add_input, add_output, contract_network, contract_pair,
                                    convert_circuit_to_network, create_ghz_preparation_circuit,
                                    create_qft_circuit, create_RQC,
                                    create_simple_preparation_circuit,
                                    full_wavefunction_contraction, load_qasm_as_circuit,
                                    network_from_dict, partition_network_on_virtual_bonds,
                                    random_contraction_plan, slice_tensor_network)
from picoquant_jl_b200.host.backends import TensorStore
from picoquant_jl_b200.host.planner import greedy_plan, sweep_plan
from picoquant_jl_b200.host.sliced import SlicedContraction, record_sliced_contraction

pytestmark = pytest.mark.gpu

DTYPES = [np.complex128, np.complex64]


def B200(dtype, **opts):
    from picoquant_jl_b200.host.b200_backend import B200Backend
    b = B200Backend(dtype)
    for k, v in opts.items():
        b.set_option(k, v)
    return b


def rand_tensor(rng, shape, dtype):
    a = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    return np.asarray(a.astype(dtype), order="F")


def oracle_amplitude(circ, plan_of, decompose, dtype=np.complex128):
    """The reference flow on the CPU oracle: the SAME network and the SAME plan
    (``plan_of`` is deterministic in the network) the device walks."""
    n = circ.n_qubits
    ob = OracleBackend(dtype)
    tn = convert_circuit_to_network(circ, ob, decompose=decompose)
    add_input(tn, "0" * n)
    add_output(tn, "0" * n)
    contract_network(tn, plan_of(tn))
    return complex(np.asarray(ob.load_tensor_data("result")).reshape(-1)[0])


def oracle_sliced_amplitude(rec, partitions, dtype=np.complex128):
    """Sum over partitions of the oracle interpreting the very command stream the
    device replays (``execute_dsl_file`` semantics, src/layer1.jl:211-315)."""
    total = 0
    for p in partitions:
        out = TensorStore()
        execute_dsl(rec.text_for(p), rec.store, dtype, output_store=out)
        total = total + out.read("result")
    return complex(np.asarray(total).reshape(-1)[0])


def close(got, ref, tol):
    """north-star bar for a scalar output tensor: relative error at 1x the tolerance."""
    return abs(complex(got) - ref) / abs(ref) < tol


# ---------------------------------------------------------------------------
# kernel level
# ---------------------------------------------------------------------------
PERMUTE_CASES = [
    ((2,) * 12, "reverse"), ((2,) * 14, "random"), ((2,) * 16, "rotate3"), ((2,) * 18, "random"),
    ((2,) * 20, "qft"), ((4, 2, 2, 8, 2, 4, 2, 2, 2, 2, 2), "random"), ((3, 5, 2, 7, 4), "random"),
    ((2, 3), "reverse"), ((1, 2, 1, 2), "random"), ((64, 64), "reverse"), ((2,) * 22, "random"),
    ((6, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2), "random"), ((2,) * 13, "identity"), ((7,), "identity"),
]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("mode", [0, 1])
def test_permute_bit_exact(dtype, mode):
    rng = np.random.default_rng(3)
    prng = random.Random(5)
    b = B200(dtype, permute=mode)
    for shape, kind in PERMUTE_CASES:
        r = len(shape)
        if kind == "reverse":
            perm = list(range(r, 0, -1))
        elif kind == "identity":
            perm = list(range(1, r + 1))
        elif kind == "rotate3":
            perm = list(range(4, r + 1)) + [1, 2, 3]
        elif kind == "qft":  # the final permute of full-wf QFT: [1,3,5,...,6,4,2]
            perm = list(range(1, r + 1, 2)) + list(range(r - (r % 2), 0, -2))
        else:
            perm = list(range(1, r + 1))
            prng.shuffle(perm)
        a = rand_tensor(rng, shape, dtype)
        b.save_tensor_data("t", a)
        b.permute_tensor("t", perm)
        got = b.load_tensor_data("t")
        ref = layer1.permute_tensor(a, perm)
        assert got.shape == ref.shape, (shape, perm)
        assert np.array_equal(got, ref), (shape, perm, kind)


def _random_contraction(rng, prng, max_labels, max_elems, extents):
    while True:
        nlab = prng.randint(1, max_labels)
        ad, ai, bd, bi = [], [], [], []
        nopen = ncon = 0
        for _ in range(nlab):
            e = prng.choice(extents)
            where = prng.choice("ABK")
            if where == "A":
                nopen += 1
                ad.append(e)
                ai.append(-nopen)
            elif where == "B":
                nopen += 1
                bd.append(e)
                bi.append(-nopen)
            else:
                ncon += 1
                ad.append(e)
                ai.append(ncon)
                bd.append(e)
                bi.append(ncon)
        pa = list(range(len(ad)))
        pb = list(range(len(bd)))
        prng.shuffle(pa)
        prng.shuffle(pb)
        ad, ai = [ad[i] for i in pa], [ai[i] for i in pa]
        bd, bi = [bd[i] for i in pb], [bi[i] for i in pb]
        if int(np.prod(ad or [1])) <= max_elems and int(np.prod(bd or [1])) <= max_elems:
            return ad, ai, bd, bi


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("opts", [dict(), dict(fused=1), dict(gemm=1), dict(gemm=3),
                                  dict(fused=1, gemm=1, permute=1)])
def test_random_contractions_match_oracle(dtype, opts):
    rng = np.random.default_rng(11)
    prng = random.Random(13)
    b = B200(dtype, **opts)
    tol = TOL[np.dtype(dtype)]
    n_done = 0
    for it in range(120):
        extents = [2] if it % 3 == 0 else ([1, 2, 4] if it % 3 == 1 else [1, 2, 3, 5])
        ad, ai, bd, bi = _random_contraction(rng, prng, 14 if it % 3 == 0 else 8, 1 << 14, extents)
        A = rand_tensor(rng, tuple(ad), dtype)
        B = rand_tensor(rng, tuple(bd), dtype)
        b.save_tensor_data("A", A)
        b.save_tensor_data("B", B)
        b.contract_tensors("A", ai, "B", bi, "C")
        got = b.load_tensor_data("C")
        ref = layer1.contract_tensors((A.astype(np.complex128), B.astype(np.complex128)), (ai, bi))
        assert got.shape == ref.shape, (ad, ai, bd, bi)
        assert rel_l2(got, ref) < tol, (ad, ai, bd, bi, rel_l2(got, ref))
        assert b.load_tensor_data("A") is None and b.load_tensor_data("B") is None
        n_done += 1
    assert n_done == 120


GEMM_SHAPES = [
    # (A dims, a_idx, B dims, b_idx): TTGT with both permutes, ragged tiles, K tails
    ((64, 8, 32), [-1, 1, -2], (16, 8, 48), [-3, 1, -4]),          # M=2048 N=768 K=8
    ((30, 7, 20), [-1, 1, -2], (7, 50), [1, -3]),                    # non-pow2, K=7
    ((2,) * 16, [-1, 1, -2, 2, -3, 3, -4, 4, -5, 5, -6, 6, -7, -8, -9, -10],
     (2,) * 12, [6, 5, 4, 3, 2, 1, -11, -12, -13, -14, -15, -16]),  # sweep-step shape M=1024 N=64 K=64
    ((128, 96), [-1, 1], (96, 80), [1, -2]),                        # plain matrix product, no permute of A
    ((96, 128), [1, -1], (80, 96), [-2, 1]),                        # both transposed
    ((4096, 2, 4), [-1, 1, 2], (4, 2, 33), [2, 1, -2]),             # N=33 ragged
    ((1, 257, 1, 19), [-1, -2, -3, 1], (19, 1, 65), [1, -4, -5]),    # extent-1 axes kept in C
]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("gemm", [0, 1])
def test_gemm_path_shapes(dtype, gemm):
    rng = np.random.default_rng(17)
    b = B200(dtype, gemm=gemm, fused=1)
    tol = TOL[np.dtype(dtype)]
    for ad, ai, bd, bi in GEMM_SHAPES:
        A = rand_tensor(rng, tuple(ad), dtype)
        B = rand_tensor(rng, tuple(bd), dtype)
        b.save_tensor_data("A", A)
        b.save_tensor_data("B", B)
        b.contract_tensors("A", ai, "B", bi, "C")
        got = b.load_tensor_data("C")
        ref = layer1.contract_tensors((A.astype(np.complex128), B.astype(np.complex128)), (ai, bi))
        assert got.shape == ref.shape
        assert rel_l2(got, ref) < tol, (ad, ai, rel_l2(got, ref))
    prof_names = None
    b.profile_enable(True)
    ad, ai, bd, bi = GEMM_SHAPES[0]
    b.save_tensor_data("A", rand_tensor(rng, tuple(ad), dtype))
    b.save_tensor_data("B", rand_tensor(rng, tuple(bd), dtype))
    b.contract_tensors("A", ai, "B", bi, "C")
    prof_names = set(b.profile_read())
    expected = "gemm_tensor" if gemm == 0 else "gemm_simt"   # DMMA (c128) / tcgen05 3xTF32 (c64)
    assert expected in prof_names, prof_names


def test_long_k_zgemm_3m():
    """Long contractions (canonical TTGT layouts behind K1): the 3M DMMA kernel (three real
    products per complex product, 16 warps, grouped rasterisation) against the oracle and
    against the four-product kernel, ragged M / N / K and several n-tile groups."""
    rng = np.random.default_rng(47)
    shapes = [
        ((200, 300), [-1, 1], (300, 130), [1, -2], dict(fused=1)),               # K=300, fused off
        ((1030, 70), [1, -1], (45, 1030), [-2, 1], dict()),                      # K > 1024: default path
        ((4, 100, 260), [1, -1, 2], (260, 4, 9, 77), [2, 1, -2, -3], dict()),    # K=1040, N=693 (11 n-tiles)
        ((2,) * 17, [1, -1, 2, -2, 3, -3, 4, -4, 5, 6, 7, 8, 9, 10, 11, -5, -6],
         (2,) * 15, [11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, -7, -8, -9, -10], dict()),  # K=2048, bit-permuted
    ]
    for ad, ai, bd, bi, opts in shapes:
        A = rand_tensor(rng, tuple(ad), np.complex128)
        B = rand_tensor(rng, tuple(bd), np.complex128)
        ref = layer1.contract_tensors((A, B), (ai, bi))
        outs = []
        for extra in (dict(), dict(zgemm_3m=1)):
            b = B200(np.complex128, **opts, **extra)
            b.save_tensor_data("A", A)
            b.save_tensor_data("B", B)
            b.profile_enable(True)
            b.contract_tensors("A", ai, "B", bi, "C")
            prof = b.profile_read()
            b.profile_enable(False)
            assert "gemm_tensor" in prof and "gemm_simt" not in prof, prof
            got = b.load_tensor_data("C")
            assert got.shape == ref.shape
            assert rel_l2(got, ref) < 1e-10, (ad, extra, rel_l2(got, ref))
            outs.append(got)
            b.close()
        assert rel_l2(outs[0], outs[1]) < 1e-13


@pytest.mark.parametrize("kfirst", [0, 1])
def test_persistent_skinny_zgemm_shapes(kfirst):
    """K <= 64, 16 < N <= 64, M >= 4 * 148 tiles of 64 rows: the persistent fused ZGEMM with B
    resident in shared memory (one CTA per SM walks the row tiles).  Ragged M / N / K, both
    gather orders, several tiles per CTA, and agreement with the tile-per-CTA kernel."""
    rng = np.random.default_rng(43)
    shapes = [
        ((2,) * 22, [-1, -2, -3, 1, 2, 3] + [-(i + 4) for i in range(10)] + [4, -14, 5, -15, 6, -16],
         (2,) * 12, [6, 5, 4, 3, 2, 1] + [-(20 + i) for i in range(6)]),       # M=2^16 N=K=64
        ((2,) * 22, [1, 2, 3] + [-(i + 1) for i in range(16)] + [4, 5, 6],
         (2,) * 11, [6, 5, 4, 3, 2, 1] + [-(20 + i) for i in range(5)]),       # low bits contracted, N=32
        ((40017, 35), [-1, 1], (35, 33), [1, -2]),                              # ragged everything
        ((7, 5, 38000), [1, 2, -1], (5, 17, 7), [2, -2, 1]),                    # K=35 first, N=17
        ((3, 50000), [1, -1], (3, 64), [1, -2]),                                # K=3
    ]
    for ad, ai, bd, bi in shapes:
        A = rand_tensor(rng, tuple(ad), np.complex128)
        B = rand_tensor(rng, tuple(bd), np.complex128)
        ref = layer1.contract_tensors((A, B), (ai, bi))
        outs = []
        # (ozaki_auto=0: this test is about the DMMA kernels; by default K >= 32, N >= 32
        # steps of this size run on the INT8 kernel, see test_default_policy_int8_kernel)
        for opts in (dict(zgemm_kfirst=kfirst, ozaki_auto=0), dict(zgemm_skinny=1)):
            b = B200(np.complex128, **opts)
            b.save_tensor_data("A", A)
            b.save_tensor_data("B", B)
            b.profile_enable(True)
            b.contract_tensors("A", ai, "B", bi, "C")
            prof = b.profile_read()
            b.profile_enable(False)
            assert set(prof) == {"gemm_tensor"}, (ad, prof)
            got = b.load_tensor_data("C")
            assert got.shape == ref.shape
            assert rel_l2(got, ref) < 1e-10, (ad, ai, opts, rel_l2(got, ref))
            outs.append(got)
            b.close()
        # same products summed in the same k order per output element => same bits
        assert rel_l2(outs[0], outs[1]) < 1e-14


@pytest.mark.parametrize("cfg", [0, 3])
def test_narrow_n_zgemm_shapes(cfg):
    """One open bond of 8..16 on the small operand with K >= 32 (a site tensor absorbed into
    the boundary): lowered to the fused DMMA GEMM with 128x8 tiles instead of the
    small-operand kernel; ragged M / N / K, both gather orders."""
    rng = np.random.default_rng(41)
    b = B200(np.complex128, zgemm_cfg=cfg, ozaki_auto=0)   # (the DMMA kernels; the default policy has its own test)
    shapes = [
        ((2,) * 19, [1, 2, 3] + [-(i + 1) for i in range(13)] + [4, 5, 6],
         (2,) * 9, [6, 5, 4, 3, 2, 1] + [-(20 + i) for i in range(3)]),        # M=2^13 N=8 K=64
        ((2,) * 18, [-(i + 1) for i in range(13)] + [1, 2, 3, 4, 5],
         (2,) * 9, [-20, 5, -21, 4, 3, -22, 2, 1, -23]),                        # N=16 K=32
        ((5000, 40), [-1, 1], (40, 12), [1, -2]),                               # ragged M, N=12
        ((33, 4100), [1, -1], (16, 33), [-2, 1]),                               # K=33, k-first A
        ((7, 4099, 5), [1, -1, 2], (9, 5, 7), [-2, 2, 1]),                      # N=9 K=35
    ]
    for ad, ai, bd, bi in shapes:
        A = rand_tensor(rng, tuple(ad), np.complex128)
        B = rand_tensor(rng, tuple(bd), np.complex128)
        b.save_tensor_data("A", A)
        b.save_tensor_data("B", B)
        b.profile_enable(True)
        b.contract_tensors("A", ai, "B", bi, "C")
        prof = b.profile_read()
        b.profile_enable(False)
        assert set(prof) == {"gemm_tensor"}, (ad, prof)
        got = b.load_tensor_data("C")
        ref = layer1.contract_tensors((A, B), (ai, bi))
        assert got.shape == ref.shape
        assert rel_l2(got, ref) < 1e-10, (ad, ai, rel_l2(got, ref))


@pytest.mark.parametrize("cfg", [0, 1, 2, 3, 4, 7])
def test_fused_ttgt_zgemm_shapes(cfg):
    """The persistent fused-TTGT ZGEMM (operands gathered inside the GEMM, no permuted
    temporaries): ragged tiles, K tails, several tiles per CTA, low-address contracted axes."""
    rng = np.random.default_rng(29)
    b = B200(np.complex128, zgemm_cfg=cfg, ozaki_auto=0)   # (the DMMA kernels; the default policy has its own test)
    shapes = list(GEMM_SHAPES) + [
        ((2,) * 20, [1, 2, 3] + [-(i + 1) for i in range(14)] + [4, 5, 6],
         (2,) * 12, [6, 5, 4, 3, 2, 1] + [-(20 + i) for i in range(6)]),       # M=2^14 N=64 K=64
        ((8, 2 ** 16, 8), [1, -1, 2], (8, 8, 40), [2, 1, -2]),                   # many tiles per CTA
        ((2 ** 17, 9), [-1, 1], (9, 70), [1, -2]),                               # KT = 1, N ragged
        ((3, 700, 5), [1, -1, 2], (5, 3, 130), [2, 1, -2]),                      # non-pow2 everything
    ]
    for ad, ai, bd, bi in shapes:
        A = rand_tensor(rng, tuple(ad), np.complex128)
        B = rand_tensor(rng, tuple(bd), np.complex128)
        b.save_tensor_data("A", A)
        b.save_tensor_data("B", B)
        b.profile_enable(True)
        b.contract_tensors("A", ai, "B", bi, "C")
        prof = b.profile_read()
        b.profile_enable(False)
        assert set(prof) == {"gemm_tensor"}, (ad, prof)   # one launch, no permute kernels
        got = b.load_tensor_data("C")
        ref = layer1.contract_tensors((A, B), (ai, bi))
        assert got.shape == ref.shape
        assert rel_l2(got, ref) < 1e-10, (ad, ai, rel_l2(got, ref))


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_large_tensor_round_trip(dtype):
    """load_tensor_data of tensors >= 64 MiB goes through the pinned two-block pipeline
    (pq_load_tensor -> d2h_pipelined): whole chunks, a ragged last chunk, bit-exact."""
    rng = np.random.default_rng(77)
    b = B200(dtype)
    item = np.dtype(dtype).itemsize
    for n in ((64 << 20) // item, (100 << 20) // item + 12345):
        x = rng.standard_normal(n).astype(np.float32).astype(dtype)
        x.imag = np.arange(n, dtype=np.float32) % 251
        b.save_tensor_data("big", x)
        got = np.asarray(b.load_tensor_data("big"))
        assert got.shape == (n,) and got.dtype == dtype
        assert np.array_equal(got, x)
        b.delete_tensor("big")
    b.close()


@pytest.mark.parametrize("tc", [0, 1])
def test_c64_small_operand_contracted_axis_fastest(tc):
    """ComplexF32 small-operand steps whose big operand has a contracted axis fastest and whose
    small operand keeps a short open bond (5..8) over 16 < K <= 64, and short contractions
    (K <= 16) with up to 64 open on the small side: k_contract_small_c64tc (16 rows per warp,
    mma.sync.m16n8k8.tf32 with 3xTF32 splitting, fragments loaded straight from the un-permuted
    operand; option small_tc = 1 switches it off: thread-per-row kernel / INT8 kernel / K1 + K3
    instead).  Small operand on either side, ragged rows / bond / K."""
    rng = np.random.default_rng(83 + tc)
    b = B200(np.complex64, small_tc=tc, ozaki_auto=0)
    shapes = [
        ((8, 5000, 8), [1, -1, 2], (8, 8, 8), [1, 2, -2]),          # K = 64, S = 8
        ((4, 4100, 8), [1, -1, 2], (4, 8, 6), [1, 2, -2]),          # K = 32, S = 6
        ((7, 4099, 5), [1, -1, 2], (7, 5, 5), [1, 2, -2]),          # K = 35, S = 5, everything ragged
        ((2,) * 18, [1, 2, 3] + [-(i + 1) for i in range(12)] + [4, 5, 6],
         (2,) * 9, [1, 2, 3, 4, 5, 6, -13, -14, -15]),              # bits, K = 64, S = 8, 2^12 rows
        ((8, 8, 8), [-1, 1, 2], (8, 4500, 8), [1, -2, 2]),          # small operand on the LEFT
        # short contraction, up to 64 open on the small side (output-bound steps)
        ((2 ** 16, 8), [-1, 1], (8, 64), [1, -2]),                  # K = 8, S = 64, rows fastest
        ((8, 2 ** 15), [1, -1], (8, 64), [1, -2]),                  # K = 8, S = 64, contracted axis fastest
        ((2, 5000, 8), [1, -1, 2], (2, 8, 32), [1, 2, -2]),         # K = 16, S = 32
        ((5, 4099), [1, -1], (5, 40), [1, -2]),                     # K = 5, S = 40, ragged
        ((3, 11, 4), [1, -1, 2], (4, 3, 6000), [2, 1, -2]),         # small (S = 11 .. K = 12) on the left: plain small kernel
        ((3, 33, 4), [1, -1, 2], (4, 3, 6000), [2, 1, -2]),         # S = 33 on the left, K = 12
    ]
    for ad, ai, bd, bi in shapes:
        A = rand_tensor(rng, tuple(ad), np.complex64)
        B = rand_tensor(rng, tuple(bd), np.complex64)
        b.save_tensor_data("A", A)
        b.save_tensor_data("B", B)
        b.profile_enable(True)
        b.contract_tensors("A", ai, "B", bi, "C")
        prof = b.profile_read()
        b.profile_enable(False)
        if tc == 0:
            assert set(prof) == {"contract_small"}, (ad, prof)
        got = b.load_tensor_data("C")
        ref = layer1.contract_tensors((A.astype(np.complex128), B.astype(np.complex128)), (ai, bi))
        assert got.shape == ref.shape
        assert rel_l2(got, ref) < 2e-6, (tc, ad, ai, rel_l2(got, ref))


@pytest.mark.parametrize("thin", [0, 1, 2, 3, 4])
def test_thin_n_zgemm(thin):
    """ComplexF64 steps with one short open bond on the small side (N <= 16, K <= 64):
    k_zgemm_thin (DMMA fragments loaded straight from the un-permuted operand; option zgemm_thin
    2 / 3 force its k-first / rows-first load order, 0 is the policy, 1 the tiled kernel it
    replaced, 4 extends the wide short-contraction case to K <= 16).  Contracted axes lowest / highest / scattered in A, N = 8 and 16, K = 32 and 64,
    ragged M, N and K, more 8-row steps than resident warps."""
    rng = np.random.default_rng(61 + thin)
    b = B200(np.complex128, zgemm_thin=thin, ozaki_auto=0)

    def case(rank_a, con, nb_open):
        ai, o, k = [], 0, 0
        for i in range(rank_a):
            if i in con:
                k += 1
                ai.append(k)
            else:
                o += 1
                ai.append(-o)
        bi = list(range(len(con), 0, -1)) + [-(o + 1 + j) for j in range(nb_open)]
        return (2,) * rank_a, ai, (2,) * (len(con) + nb_open), bi

    shapes = [
        case(20, [0, 1, 2, 15, 17, 19], 3),      # contracted axes lowest: k-first policy, K = 64, N = 8
        case(20, [3, 4, 5, 14, 16, 18], 3),      # open axes lowest
        case(19, [0, 1, 16, 17, 18], 4),         # K = 32, N = 16
        case(22, [1, 3, 5, 7, 9, 11], 3),        # 2^16 rows: several steps per warp
        ((4099, 37), [-1, 1], (37, 11), [1, -2]),                 # ragged M, N, K (rows fastest)
        ((53, 5001), [1, -1], (53, 13), [1, -2]),                 # ragged, contracted axis fastest
        ((8, 4100, 8), [1, -1, 2], (8, 8, 9), [2, 1, -2]),        # two contracted axes around the open one
        # short contraction, wide small side (policy: K <= 8; option 4: K <= 16)
        case(19, [16, 17, 18], 6),               # M = 2^16, N = 64, K = 8, contracted axes highest
        case(19, [0, 1, 2], 5),                  # N = 32, contracted axes lowest
        ((4099, 7), [-1, 1], (7, 37), [1, -2]),                   # ragged
        ((13, 5001), [1, -1], (13, 50), [1, -2]),                 # K = 13 first
    ]
    for ad, ai, bd, bi in shapes:
        A = rand_tensor(rng, tuple(ad), np.complex128)
        B = rand_tensor(rng, tuple(bd), np.complex128)
        b.save_tensor_data("A", A)
        b.save_tensor_data("B", B)
        b.profile_enable(True)
        b.contract_tensors("A", ai, "B", bi, "C")
        prof = b.profile_read()
        b.profile_enable(False)
        assert set(prof) == {"gemm_tensor"}, (ad, prof)
        got = b.load_tensor_data("C")
        ref = layer1.contract_tensors((A, B), (ai, bi))
        assert got.shape == ref.shape
        assert rel_l2(got, ref) < 1e-13, (thin, ad, ai, rel_l2(got, ref))


@pytest.mark.parametrize("K", [8, 40, 255, 256, 257, 513, 4096])
def test_tcgen05_cgemm_accuracy_over_k(K):
    """ComplexF32 GEMM on tcgen05 with 3xTF32 splitting and K-chunked fp32 folding: the
    rel-L2 error against the ComplexF64 oracle must stay below 1e-5 for short and long
    contractions, including the chunk boundaries (KCHUNK = 256)."""
    rng = np.random.default_rng(31 + K)
    b = B200(np.complex64, fused=1)
    M, N = 384 + 5, 96 + 3
    A = rand_tensor(rng, (M, K), np.complex64)
    B = rand_tensor(rng, (K, N), np.complex64)
    b.save_tensor_data("A", A)
    b.save_tensor_data("B", B)
    b.profile_enable(True)
    b.contract_tensors("A", [-1, 1], "B", [1, -2], "C")
    prof = b.profile_read()
    b.profile_enable(False)
    assert "gemm_tensor" in prof and "gemm_simt" not in prof
    got = b.load_tensor_data("C")
    ref = A.astype(np.complex128) @ B.astype(np.complex128)
    assert rel_l2(got, ref) < 1e-5, (K, rel_l2(got, ref))


@pytest.mark.parametrize("dtype", DTYPES)
def test_plan_independence_rqc_4x5(dtype):
    """Size-independent property for the RQC amplitude configs: two different contraction
    plans (greedy on the undecomposed network -- GEMM heavy, K up to 2^10 -- and the sweep
    plan on the decomposed one) and the sliced replay must agree on the amplitude."""
    tol = TOL[np.dtype(dtype)]
    circ = create_RQC(4, 5, 16, seed=21)
    n = circ.n_qubits
    vals = []
    for decompose in (False, True):
        def plan_of(tn):
            return sweep_plan(tn, 4, 5) if decompose else greedy_plan(tn)
        b = B200(dtype)
        tn = convert_circuit_to_network(circ, b, decompose=decompose)
        add_input(tn, "0" * n)
        add_output(tn, "0" * n)
        contract_network(tn, plan_of(tn))
        vals.append(complex(b.load_tensor_data("result")))
        # parity: the ComplexF64 oracle walking the same plan, at 1x the tolerance
        ref = oracle_amplitude(circ, plan_of, decompose)
        assert close(vals[-1], ref, tol), (decompose, vals[-1], ref)
    # second, looser sanity check: the dense state-vector simulation (a different algorithm)
    sim = circ.simulate()[0]
    for v in vals:
        assert abs(v - sim) / abs(sim) < 100 * tol, (vals, sim)
    assert abs(vals[0] - vals[1]) / abs(sim) < 2 * tol


@pytest.mark.parametrize("dtype", DTYPES)
def test_default_policy_int8_kernel(dtype):
    """Default kernel choice (option ozaki_auto, csrc/common.h: ozaki_t_preferred): the skinny
    GEMM-shaped sweep steps run on the INT8 tensor-core kernel k_ozaki_t with the gather fused
    (one launch, class gemm_int8, no permute kernels) and meet the north-star tolerance at 1x;
    shapes outside the policy keep their DMMA / fused small-operand kernels.  Ragged M / N / K,
    contracted axes scattered over A, K first, narrow N, odd M (scalar store path)."""
    rng = np.random.default_rng(77)
    tol = TOL[np.dtype(dtype)]
    c128 = np.dtype(dtype) == np.complex128
    cases = [
        # M = 2^16, N = K = 64, contracted axes scattered
        ((2,) * 22, [-1, -2, -3, 1, 2, 3] + [-(i + 4) for i in range(10)] + [4, -14, 5, -15, 6, -16],
         (2,) * 12, [6, 5, 4, 3, 2, 1] + [-(20 + i) for i in range(6)], True),
        # low bits contracted, N = 32
        ((2,) * 22, [1, 2, 3] + [-(i + 1) for i in range(16)] + [4, 5, 6],
         (2,) * 11, [6, 5, 4, 3, 2, 1] + [-(20 + i) for i in range(5)], True),
        ((40017, 35), [-1, 1], (35, 33), [1, -2], True),          # ragged everything, odd M
        ((7, 5, 38000), [1, 2, -1], (5, 17, 7), [2, -2, 1], not c128),   # K = 35 first, N = 17: c64 only
        ((3, 50000), [1, -1], (3, 64), [1, -2], False),           # K = 3: output-bound; DMMA kernel (c128) / small-operand tensor-core kernel (c64)
        ((2000, 64), [-1, 1], (64, 64), [1, -2], False),          # M < 4096: outside the envelope
    ]
    for ad, ai, bd, bi, int8 in cases:
        A = rand_tensor(rng, tuple(ad), dtype)
        B = rand_tensor(rng, tuple(bd), dtype)
        ref = layer1.contract_tensors((A.astype(np.complex128), B.astype(np.complex128)), (ai, bi))
        b = B200(dtype)
        b.save_tensor_data("A", A)
        b.save_tensor_data("B", B)
        b.profile_enable(True)
        b.contract_tensors("A", ai, "B", bi, "C")
        prof = b.profile_read()
        b.profile_enable(False)
        if int8:
            assert set(prof) == {"gemm_int8"}, (ad, prof)
        else:
            assert "gemm_int8" not in prof, (ad, prof)
        got = b.load_tensor_data("C")
        assert got.shape == ref.shape
        assert rel_l2(got, ref) < tol, (ad, ai, rel_l2(got, ref))
        assert b.microbench("ozaki_t_debug") == 0   # no mbarrier watchdog event
        b.close()


def test_int8_kernel_special_values():
    """k_ozaki_t on rows / columns the slicing treats specially: an all-zero row gives exact
    zeros, rows 1e+-140 apart keep their relative accuracy (one power-of-two scale per row and
    column), a row holding an Inf or a NaN gives NaN in that row only (as DMMA would)."""
    rng = np.random.default_rng(5)
    M, N, K = 8192, 64, 64
    A = rand_tensor(rng, (M, K), np.complex128)
    B = rand_tensor(rng, (K, N), np.complex128)
    A[3, :] = 0
    A[10, :] *= 1e140
    A[11, :] *= 1e-140
    B[:, 5] *= 1e-120
    B[:, 6] *= 1e120
    A[20, 7] = np.inf
    A[21, 9] = np.nan
    b = B200(np.complex128)
    b.save_tensor_data("A", A)
    b.save_tensor_data("B", B)
    b.profile_enable(True)
    b.contract_tensors("A", [-1, 1], "B", [1, -2], "C")
    assert set(b.profile_read()) == {"gemm_int8"}
    got = np.asarray(b.load_tensor_data("C"))
    with np.errstate(all="ignore"):
        ref = A @ B
    assert np.all(got[3] == 0)
    assert np.all(np.isnan(got[20])) and np.all(np.isnan(got[21]))
    ok = np.ones(M, bool)
    ok[[20, 21]] = False
    # every row and column judged at its own scale
    err = np.abs(got[ok] - ref[ok]) / (np.linalg.norm(A[ok], axis=1)[:, None] * np.linalg.norm(B, axis=0)[None, :] + 1e-300)
    assert err.max() < 1e-11, err.max()
    b.close()


@pytest.mark.parametrize("dtype", DTYPES)
def test_fused_kernel_classes_and_dot(dtype):
    """Gate application (small right), cap contraction (small left) and the final
    inner product (dot) must run as single fused launches, and match the oracle."""
    rng = np.random.default_rng(19)
    b = B200(dtype)
    tol = TOL[np.dtype(dtype)]
    cases = [
        ((2,) * 18, [-(i + 1) if i not in (4, 11) else (1 if i == 4 else 2) for i in range(18)],
         (2, 2, 2, 2), [2, 1, -30, -31], "contract_small"),
        ((2, 2), [-1, 1], (2,) * 16, [-(i + 2) if i != 7 else 1 for i in range(16)], "contract_small"),
        ((2,) * 17, list(range(1, 18)), (2,) * 17, list(range(17, 0, -1)), "contract_dot"),
        ((2,) * 15 + (2,), list(range(1, 16)) + [-1], (2,) * 15, list(range(15, 0, -1)), "contract_dot"),
    ]
    for ad, ai, bd, bi, cls in cases:
        # renumber open labels to be distinct negatives
        A = rand_tensor(rng, tuple(ad), dtype)
        B = rand_tensor(rng, tuple(bd), dtype)
        b.save_tensor_data("A", A)
        b.save_tensor_data("B", B)
        b.profile_enable(True)
        b.contract_tensors("A", ai, "B", bi, "C")
        prof = b.profile_read()
        b.profile_enable(False)
        assert set(prof) == {cls}, prof
        got = b.load_tensor_data("C")
        ref = layer1.contract_tensors((A.astype(np.complex128), B.astype(np.complex128)), (ai, bi))
        assert got.shape == ref.shape
        assert rel_l2(got, ref) < tol, (cls, rel_l2(got, ref))


@pytest.mark.parametrize("dtype", DTYPES)
def test_backend_semantics(dtype):
    """interactive.jl semantics: conversion on save, None for missing labels,
    KeyError on missing operands, alias on save_output, view keeps the axis,
    reshape groups, delete of a missing label is fine."""
    rng = np.random.default_rng(23)
    b = B200(dtype)
    assert b.load_tensor_data("nope") is None
    b.delete_tensor("nope")
    with pytest.raises(KeyError):
        b.contract_tensors("x", [1], "y", [1], "z")
    # real float64 input is converted to the backend's complex type
    b.save_tensor_data("cap", np.array([1.0, 0.0]))
    got = b.load_tensor_data("cap")
    assert got.dtype == np.dtype(dtype) and np.array_equal(got, np.array([1, 0], dtype=dtype))
    a = rand_tensor(rng, (2, 3, 4, 5), np.complex128)
    b.save_tensor_data("a", a)
    assert rel_l2(b.load_tensor_data("a"), a.astype(dtype)) == 0.0
    b.view_tensor("v", "a", 3, range(2, 4))
    assert np.array_equal(b.load_tensor_data("v"), a.astype(dtype)[:, :, 1:3, :])
    b.view_tensor("v1", "a", 1, range(2, 3))
    assert b.load_tensor_data("v1").shape == (1, 3, 4, 5)
    b.view_tensor("v2", "a", 4, [1, 3, 4])
    assert np.array_equal(b.load_tensor_data("v2"), a.astype(dtype)[:, :, :, [0, 2, 3]])
    b.save_output("a", "result")
    b.permute_tensor("a", [2, 1, 3, 4])                     # rebinding must not touch the alias
    assert np.array_equal(b.load_tensor_data("result"), a.astype(dtype))
    assert b.load_tensor_data("a").shape == (3, 2, 4, 5)
    b.reshape_tensor("a", [[1, 2], [3], [4]])
    assert b.load_tensor_data("a").shape == (6, 4, 5)
    b.reshape_tensor("a", [[1, 2, 3]])
    assert b.load_tensor_data("a").shape == (120,)
    with pytest.raises(ValueError):
        b.save_tensor_data("p", rand_tensor(rng, (2, 3), dtype))
        b.save_tensor_data("q", rand_tensor(rng, (4, 2), dtype))
        b.contract_tensors("p", [-1, 1], "q", [1, -2], "r")
    # scalars: rank-0 result and rank-0 operand
    b.save_tensor_data("s1", rand_tensor(rng, (2,), dtype))
    b.save_tensor_data("s2", rand_tensor(rng, (2,), dtype))
    s1, s2 = b.load_tensor_data("s1"), b.load_tensor_data("s2")
    b.contract_tensors("s1", [1], "s2", [1], "s")
    assert b.load_tensor_data("s").shape == ()
    assert abs(b.load_tensor_data("s") - np.sum(s1.astype(np.complex128) * s2)) < 1e-5
    b.save_tensor_data("w", rand_tensor(rng, (3,), dtype))
    w = b.load_tensor_data("w")
    sval = b.load_tensor_data("s")
    b.contract_tensors("s", [], "w", [-1], "sw")
    assert rel_l2(b.load_tensor_data("sw"), sval * w) < TOL[np.dtype(dtype)]
    b.accumulate("acc", "sw")
    b.accumulate("acc", "sw")
    assert rel_l2(b.load_tensor_data("acc"), 2 * sval * w) < TOL[np.dtype(dtype)]


# ---------------------------------------------------------------------------
# the reference's own tests, on the device backend
# ---------------------------------------------------------------------------
def test_reference_stored_contraction_golden_on_gpu():
    d = load_golden_json("ghz_3.json")
    g = load_golden_json("ghz_3_contracted.json")
    b = B200(np.complex128)
    tn = network_from_dict(d, b)
    for k, v in d["nodes"].items():
        data = np.array(v["data_re"]) + 1j * np.array(v["data_im"])
        b.save_tensor_data(k, np.reshape(data, v["data_dims"], order="F"))
    for edge in load_golden_json("ghz_3_plan.json"):
        contract_pair(tn, edge)
    (label, gnode), = g["nodes"].items()
    out = b.load_tensor_data(label)
    ref = np.array(gnode["data_re"]) + 1j * np.array(gnode["data_im"])
    assert list(out.shape) == gnode["data_dims"]
    assert rel_l2(out.ravel(order="F"), ref) < 1e-15


@pytest.mark.parametrize("dtype", DTYPES)
def test_reference_known_answers_on_gpu(dtype):
    tol = TOL[np.dtype(dtype)]
    # metrics golden + GHZ-3 (test/layer2_tests.jl:106-143, layer1_tests.jl:11-46)
    b = B200(dtype)
    c = Circuit(3).h(0).cx(0, 1).cx(0, 1).cx(0, 2)
    tn = convert_circuit_to_network(c, b)
    add_input(tn, "000")
    full_wavefunction_contraction(tn, "vector")
    assert b.metrics.as_tuple() == (8, 44, 124)
    cnt = b.counters()
    assert cnt["n_contract"] == 6 and cnt["macs"] == 124 and cnt["kernel_launches"] >= 6
    ghz = load_qasm_as_circuit(golden_qasm("ghz_3.qasm"))
    for seed in range(3):
        bb = B200(dtype)
        tn = convert_circuit_to_network(ghz, bb)
        add_input(tn, "000")
        add_output(tn, "000")
        contract_network(tn, random_contraction_plan(tn, random.Random(seed)))
        res = bb.load_tensor_data("result")
        assert res.shape == () and abs(res - 1 / np.sqrt(2)) < 10 * tol
    # disjoint pieces (layer2_tests.jl:69-103)
    bb = B200(dtype)
    tn = convert_circuit_to_network(Circuit(2).h(0).h(1), bb)
    add_input(tn, "00")
    contract_network(tn, random_contraction_plan(tn, random.Random(1)), "vector")
    res = bb.load_tensor_data("result")
    assert res.ndim == 1 and abs(res.real[0] - 0.5) < 10 * tol
    # GHZ-5 (layer2_tests.jl:308-328)
    psi = statevector(create_ghz_preparation_circuit(5), B200(dtype))
    ref = np.zeros(32, dtype=np.complex128)
    ref[[0, -1]] = 1 / np.sqrt(2)
    assert rel_l2(psi, ref) < tol
    # decomposed gate re-contraction (layer3_tests.jl:57-71)
    rng = np.random.default_rng(7)
    gate = rand_tensor(rng, (2, 2, 2, 2), np.complex128)
    bb = B200(dtype)
    tn = TensorNetworkCircuit(2, bb)
    out = contract_pair(tn, *add_gate(tn, gate, [1, 2], decompose=True))
    assert rel_l2(np.transpose(bb.load_tensor_data(out), (0, 2, 1, 3)), gate) < 10 * tol


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name,n", [("qft_3.qasm", 3), ("qft_5.qasm", 5), ("qft_10.qasm", 10)])
def test_qft_fixtures_on_gpu(dtype, name, n):
    """configs 1-2: closed form for |0..0> and parity with the oracle for a
    non-trivial input ('1', '+', '-' caps)."""
    tol = TOL[np.dtype(dtype)]
    circ = load_qasm_as_circuit(golden_qasm(name))
    psi = statevector(circ, B200(dtype))
    assert rel_l2(psi, np.full(2 ** n, 2 ** (-n / 2))) < tol
    cfg = ("1+-0" * n)[:n]
    got = statevector(circ, B200(dtype), input_config=cfg)
    ref = statevector(circ, OracleBackend(np.complex128), input_config=cfg)
    assert rel_l2(got, ref) < tol


@pytest.mark.parametrize("dtype", DTYPES)
def test_qft_against_inverse_fft_on_gpu(dtype):
    """test/algorithms_tests.jl:39-82 with n = 8 and 14."""
    tol = TOL[np.dtype(dtype)]
    for n in (8, 14):
        prep = create_simple_preparation_circuit(n, 3, 43)
        full = prep.compose(create_qft_circuit(n))
        psi_in = statevector(prep, OracleBackend(np.complex128))
        ref = np.fft.ifft(switch_endianness(psi_in))
        ref /= np.linalg.norm(ref)
        psi = switch_endianness(statevector(full, B200(dtype)))
        assert abs(abs(np.vdot(psi, ref)) - 1.0) < 10 * tol
        oracle = statevector(full, OracleBackend(np.complex128))
        assert rel_l2(switch_endianness(psi), oracle) < tol


@pytest.mark.parametrize("dtype", DTYPES)
def test_slicing_identity_on_gpu(dtype):
    """test/layer2_tests.jl:419-455 through the backend calls (view_tensor!)."""
    tol = TOL[np.dtype(dtype)]
    n = 4
    circ = create_simple_preparation_circuit(n, 2, 5).compose(create_qft_circuit(n))
    wf = statevector(circ, OracleBackend(np.complex128), decompose=True)
    for P in (4, 8):
        total = np.zeros(2 ** n, dtype=np.complex128)
        for p in range(1, P + 1):
            b = B200(dtype)
            tn = convert_circuit_to_network(circ, b, decompose=True)
            add_input(tn, "0" * n)
            labels, values = partition_network_on_virtual_bonds(tn, P, p)
            slice_tensor_network(tn, labels, values)
            full_wavefunction_contraction(tn, "vector")
            total += b.load_tensor_data("result")
        assert rel_l2(total, wf) < tol


# ---------------------------------------------------------------------------
# .tl programs (execute_dsl_file on the device) and sliced replay
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("graph", [0, 1, 2])
def test_program_matches_dsl_interpreter(dtype, graph):
    tol = TOL[np.dtype(dtype)]
    circ = create_RQC(3, 3, 8, seed=4)
    dsl = DSLBackend()
    tn = convert_circuit_to_network(circ, dsl, decompose=True)
    add_input(tn, "0" * 9)
    full_wavefunction_contraction(tn, "vector")
    out = TensorStore()
    execute_dsl(dsl.text(), dsl.store, np.complex128, output_store=out)
    ref = out.read("result")
    assert rel_l2(ref, circ.simulate()) < 1e-12
    b = B200(dtype, graph=graph)
    for key, arr in dsl.store.data.items():
        b.save_tensor_data(key, arr)
    prog = b.compile_program(dsl.text())
    assert prog.num_views == 0 and prog.launches > 0 and prog.arena_bytes > 0
    for _ in range(3):  # replays must be idempotent (leaves are not consumed)
        prog.run()
        assert rel_l2(b.load_tensor_data("result"), ref) < tol
    assert b.counters()["kernel_launches"] >= 3 * prog.launches


@pytest.mark.parametrize("graph", [0, 2])
def test_large_program_dag_vs_chain(graph):
    """A 5x5 depth-16 sliced amplitude (GEMM steps, arena reuse, ~500 launches): the
    multi-stream DAG capture must give the same amplitude as the single-chain graph and
    as the oracle's interpreter."""
    circ = create_RQC(5, 5, 16, seed=3)
    n = circ.n_qubits

    def plan_fn(tn, sliced):
        return sweep_plan(tn, 5, 5, sliced_bonds=sliced)

    P = 8
    rec = record_sliced_contraction(circ, P, 1, plan_fn=plan_fn, output_config="0" * n)
    ref = 0
    for p in range(1, P + 1):
        out = TensorStore()
        execute_dsl(rec.text_for(p), rec.store, np.complex128, output_store=out)
        ref = ref + out.read("result")
    b = B200(np.complex128, graph=graph)
    sc = SlicedContraction(b, rec)
    assert sc.program.macs_invariant + sc.program.macs_dependent == sc.program.macs
    assert sc.program.macs_invariant > 0 and sc.program.macs_dependent > 0
    for hoist in (False, True, True, False):
        # hoisting runs the slice-invariant steps once per call instead of once per slice
        b.delete_tensor("partial_sum")
        b.reset_counters()
        sc.run(range(1, P + 1), hoist=hoist)
        got = sc.result()
        assert abs(got - ref) / abs(ref) < 1e-10, (graph, hoist, got, ref)
        macs = b.counters()["macs"]
        expect = (sc.program.macs_invariant + P * sc.program.macs_dependent) if hoist \
            else P * sc.program.macs
        assert macs == expect, (hoist, macs, expect)


@pytest.mark.parametrize("dtype", DTYPES)
def test_chains_of_tiny_contractions(dtype):
    """Compiled programs batch chains of tiny contractions (the world-lines of a slice) into a
    few launches, one CTA per chain.  Same amplitude as with one launch per contraction and as
    the oracle, far fewer launches per replay, with and without hoisting and with lanes."""
    tol = TOL[np.dtype(dtype)]
    circ = create_RQC(4, 5, 14, seed=5)
    n = circ.n_qubits
    P = 8

    def plan_fn(tn, sliced):
        return sweep_plan(tn, 4, 5, sliced_bonds=sliced)

    rec = record_sliced_contraction(circ, P, 1, plan_fn=plan_fn, output_config="0" * n)
    ref = oracle_sliced_amplitude(rec, range(1, P + 1))   # same streams, ComplexF64 oracle
    assert abs(ref - circ.simulate()[0]) / abs(ref) < 1e-10
    results, launches = {}, {}
    for chain in (0, 1):
        b = B200(dtype, chain=chain)
        sc = SlicedContraction(b, rec)
        for hoist, lanes in ((False, 1), (True, 1), (False, 3), (True, 2)):
            b.delete_tensor("partial_sum")
            b.reset_counters()
            sc.run(range(1, P + 1), hoist=hoist, lanes=lanes)
            got = sc.result()
            assert close(got, ref, tol), (chain, hoist, lanes, got, ref)
            results[(chain, hoist, lanes)] = got.copy()
            if not hoist and lanes == 1:
                launches[chain] = b.counters()["kernel_launches"] // P
            assert b.counters()["macs"] == ((sc.program.macs_invariant + P * sc.program.macs_dependent)
                                            if hoist else P * sc.program.macs)
        # every way of running the same program gives the same bits
        base = results[(chain, False, 1)]
        for key, v in results.items():
            if key[0] == chain:
                assert v.tobytes() == base.tobytes(), key
        b.close()
    assert launches[0] < launches[1] / 3, launches
    assert abs(results[(0, False, 1)] - results[(1, False, 1)]) / abs(ref) < tol


@pytest.mark.parametrize("dtype", DTYPES)
def test_sliced_program_replay(dtype):
    """One compiled plan replayed for every partition; the stream for partition p
    equals the one the host mirror emits for p (only view indices differ)."""
    tol = TOL[np.dtype(dtype)]
    circ = create_RQC(3, 4, 10, seed=2)
    n = circ.n_qubits

    def plan_fn(tn, sliced):
        return sweep_plan(tn, 3, 4, sliced_bonds=sliced)

    for P in (4, 16):
        rec = record_sliced_contraction(circ, P, 1, plan_fn=plan_fn, output_config="0" * n)
        other = record_sliced_contraction(circ, P, P, plan_fn=plan_fn, output_config="0" * n)
        assert rec.text_for(P) == other.text and rec.text_for(1) == rec.text
        b = B200(dtype)
        sc = SlicedContraction(b, rec)
        sc.run(range(1, P + 1))
        got = sc.result()
        ref = oracle_sliced_amplitude(rec, range(1, P + 1))
        assert abs(ref - circ.simulate()[0]) / abs(ref) < 1e-10
        assert got.shape == ()
        assert close(got, ref, tol), (P, got, ref)
        # re-uploading the gate tensors (same shapes) updates the bound buffers in place
        sc.upload()
        b.delete_tensor("partial_sum")
        sc.run(range(1, P + 1))
        assert close(sc.result(), ref, tol)
    # rebinding a leaf to a different shape invalidates the program (no stale reads)
    from picoquant_jl_b200.host.b200_backend import B200Error
    first_leaf = rec.text.split()[2]
    b.save_tensor_data(first_leaf, np.zeros((3, 3), dtype=dtype))
    with pytest.raises(B200Error):
        sc.program.run(rec.view_starts(1), None)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("hoist", [False, True])
def test_slice_lanes_bit_identical(dtype, hoist):
    """pq_program_run_slices keeps several partitions in flight on private arenas / streams;
    the partial sums are added in partition order, so the amplitude must be BIT-identical to
    the one-partition-at-a-time loop, for every lane count, with and without hoisting."""
    circ = create_RQC(3, 4, 10, seed=2)
    n = circ.n_qubits
    P = 16

    def plan_fn(tn, sliced):
        return sweep_plan(tn, 3, 4, sliced_bonds=sliced)

    rec = record_sliced_contraction(circ, P, 1, plan_fn=plan_fn, output_config="0" * n)
    b = B200(dtype)
    sc = SlicedContraction(b, rec)
    sc.run(range(1, P + 1), hoist=hoist)
    base = sc.result().copy()
    ref = oracle_sliced_amplitude(rec, range(1, P + 1))
    assert close(base, ref, TOL[np.dtype(dtype)]), (base, ref)
    for lanes in (2, 3, 8):
        for rep in range(2):   # second repetition reuses the lanes' graphs and arenas
            b.delete_tensor("partial_sum")
            b.reset_counters()
            sc.run(range(1, P + 1), hoist=hoist, lanes=lanes)
            got = sc.result()
            assert got.tobytes() == base.tobytes(), (lanes, rep, got, base)
            if not hoist:
                assert b.counters()["macs"] == P * sc.program.macs
    # a partial range (one rank's share) and a following single run still work
    b.delete_tensor("partial_sum")
    sc.run(range(5, 12), hoist=hoist, lanes=4)
    sc.run([12], hoist=hoist)
    part = sc.result().copy()
    b.delete_tensor("partial_sum")
    sc.run(range(5, 13), hoist=hoist)
    assert part.tobytes() == sc.result().tobytes()


@pytest.mark.parametrize("dtype", DTYPES)
def test_rqc_amplitude_plans(dtype):
    """config-4 shape at test size: single amplitude of a 4x4 depth-16 RQC with the
    greedy plan (undecomposed network, GEMM-heavy) and the sweep plan (decomposed)."""
    tol = TOL[np.dtype(dtype)]
    circ = create_RQC(4, 4, 16, seed=9)
    sim = circ.simulate()[0]
    for decompose in (False, True):
        def plan_of(tn):
            return sweep_plan(tn, 4, 4) if decompose else greedy_plan(tn)
        b = B200(dtype)
        tn = convert_circuit_to_network(circ, b, decompose=decompose)
        add_input(tn, "0" * 16)
        add_output(tn, "0" * 16)
        contract_network(tn, plan_of(tn))
        got = b.load_tensor_data("result")
        ref = oracle_amplitude(circ, plan_of, decompose)   # same plan, ComplexF64 oracle
        assert close(got, ref, tol), (decompose, got, ref)
        assert abs(got - sim) / abs(sim) < 50 * tol, (decompose, got, sim)


def test_config4_rqc_6x6_d20_amplitude():
    """BASELINE config 4 at full size: <0..0|C|0..0> of the 6x6 depth-20 RQC, undecomposed
    network, noise-free greedy pair plan (489 contractions, 2^37.1 complex MACs, top step
    M=N=2^13 K=2^11, largest intermediate 2^26 elements), as one compiled program on the GPU
    against the NumPy/OpenBLAS oracle walking the same plan on the host cores.
    ComplexF64: relative error <= 1e-10.  ComplexF32 (tcgen05 3xTF32): <= 1e-5 against the
    ComplexF64 oracle (no multiplier, no escape clause)."""
    n = 36
    circ = create_RQC(6, 6, 20, seed=0)

    def network(backend):
        tn = convert_circuit_to_network(circ, backend, decompose=False)
        add_input(tn, "0" * n)
        add_output(tn, "0" * n)
        return tn

    dsl = DSLBackend()
    tn = network(dsl)
    plan = greedy_plan(tn)
    contract_network(tn, plan, "")
    refs = {}
    for dtype in DTYPES:
        ob = OracleBackend(dtype)
        contract_network(network(ob), plan, "")
        refs[np.dtype(dtype)] = complex(np.asarray(ob.load_tensor_data("result")).reshape(-1)[0])
    ref64 = refs[np.dtype(np.complex128)]
    for dtype in DTYPES:
        b = B200(dtype)
        for key, arr in dsl.store.data.items():
            b.save_tensor_data(key, arr)
        prog = b.compile_program(dsl.text())
        assert prog.macs == 147285958496
        prog.run()
        got = complex(np.asarray(b.load_tensor_data("result")).reshape(-1)[0])
        err = abs(got - ref64) / abs(ref64)
        assert err < TOL[np.dtype(dtype)], (got, ref64, err,
                                             abs(refs[np.dtype(np.complex64)] - ref64) / abs(ref64))
        prog.close()
        b.close()


# ---------------------------------------------------------------------------
# BASELINE sizes through size-independent properties
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
def test_qft_20_closed_form(dtype):
    """Closed form of the reference's QFT circuit for a basis input: in big-endian
    indexing the circuit is the normalised inverse DFT (test/algorithms_tests.jl:39-82),
    so amp_BE[k] = 2^{-n/2} exp(2 pi i x_BE k / 2^n)."""
    n = 20
    tol = TOL[np.dtype(dtype)]
    circ = create_qft_circuit(n)
    x_le = 0b10110011100011110001
    cfg = "".join("1" if (x_le >> q) & 1 else "0" for q in range(n))
    x_be = int(cfg, 2)  # qubit 0 becomes the most significant bit
    psi = statevector(circ, B200(dtype), input_config=cfg)
    k = np.arange(2 ** n, dtype=np.int64)
    ref_be = np.exp(2j * np.pi * ((x_be * k) % (2 ** n)) / 2 ** n) * 2 ** (-n / 2)
    assert rel_l2(switch_endianness(psi), ref_be) < tol


@pytest.mark.parametrize("dtype", DTYPES)
def test_qft_26_closed_form(dtype):
    """config 3 at full size (2^26 amplitudes) with a NON-trivial basis input, so every cu1
    phase matters: the closed form of test/algorithms_tests.jl:39-82 (normalised inverse DFT
    in big-endian indexing), rel-L2 at 1x the north-star tolerance, both element types."""
    n = 26
    tol = TOL[np.dtype(dtype)]
    circ = create_qft_circuit(n)
    x_le = 0b10110011100011110001101101
    cfg = "".join("1" if (x_le >> q) & 1 else "0" for q in range(n))
    x_be = int(cfg, 2)
    b = B200(dtype)
    psi = statevector(circ, b, input_config=cfg)
    assert psi.shape == (2 ** n,) and b.counters()["n_contract"] == 389
    b.close()
    psi_be = switch_endianness(psi)
    del psi
    num = den = 0.0
    step = 2 ** 22   # closed form in blocks (keeps the host footprint small)
    for k0 in range(0, 2 ** n, step):
        k = np.arange(k0, k0 + step, dtype=np.int64)
        ref = np.exp(2j * np.pi * ((x_be * k) % (2 ** n)) / 2 ** n) * 2 ** (-n / 2)
        d = psi_be[k0:k0 + step].astype(np.complex128) - ref
        num += float(np.vdot(d, d).real)
        den += float(np.vdot(ref, ref).real)
    assert np.sqrt(num / den) < tol, np.sqrt(num / den)


@pytest.mark.parametrize("dtype", DTYPES)
def test_config5_rqc_7x7_d24_slices_vs_oracle(dtype):
    """BASELINE config 5 at full size (the bench workload): 7x7 depth-24 RQC, decomposed,
    P = 64 slices, sweep plan.  The device partial sum of four slices (first, two middle,
    last; two lanes, like the timed loop) against the ComplexF64 oracle interpreting the
    same command streams, at 1x the tolerance (test/layer2_tests.jl:419-455 at scale)."""
    tol = TOL[np.dtype(dtype)]
    circ = create_RQC(7, 7, 24, seed=0)
    n = circ.n_qubits

    def plan_fn(tn, sliced):
        return sweep_plan(tn, 7, 7, sliced_bonds=sliced)

    rec = record_sliced_contraction(circ, 64, 1, plan_fn=plan_fn, output_config="0" * n)
    sample = [1, 22, 43, 64]
    ref = oracle_sliced_amplitude(rec, sample)
    b = B200(dtype)
    sc = SlicedContraction(b, rec)
    sc.run(sample, lanes=2)
    got = sc.result()
    assert close(got, ref, tol), (got, ref, abs(complex(got) - ref) / abs(ref))
    b.close()


def test_qft_26_uniform_c64():
    """config 3 at full size (2^26 amplitudes, ComplexF32): uniform superposition."""
    n = 26
    b = B200(np.complex64)
    psi = statevector(create_qft_circuit(n), b)
    assert psi.shape == (2 ** n,)
    assert abs(np.linalg.norm(psi.astype(np.complex128)) - 1.0) < 1e-4
    assert np.max(np.abs(psi - 2 ** (-n / 2))) < 1e-5 * 2 ** (-n / 2) * 50
    assert b.counters()["n_contract"] == 389


@pytest.mark.parametrize("dtype", DTYPES)
def test_batched_save_matches_single_saves(dtype):
    """pq_save_tensors (one staging block, one H2D, one scatter launch) must leave the store in
    exactly the state n pq_save_tensor calls do: bit-identical data, conversion to the backend
    dtype, in-place update of bound leaves (a compiled program sees the new values), rebinding
    on a size change, rank-0 tensors."""
    rng = np.random.default_rng(77)
    items = [("g%d" % i, rand_tensor(rng, shp, np.complex128))
             for i, shp in enumerate([(2, 2), (2, 2, 2, 2), (2,), (4, 2, 2), (3, 5), (1,), (2, 2, 4)])]
    items.append(("real64", np.asfortranarray(rng.standard_normal((2, 3)))))
    items.append(("real32", np.asfortranarray(rng.standard_normal((4,)).astype(np.float32))))
    items.append(("scalar", np.array(1.5 - 2j)))
    items.append(("c64src", rand_tensor(rng, (2, 2), np.complex64)))
    a, b = B200(dtype), B200(dtype)
    for label, arr in items:
        a.save_tensor_data(label, arr)
    b.save_tensors(items)
    for label, arr in items:
        x, y = a.load_tensor_data(label), b.load_tensor_data(label)
        assert x.shape == y.shape == np.asarray(arr).shape
        assert x.tobytes() == y.tobytes(), label
    # a program bound to the leaves sees a batched in-place re-save
    text = "tensor A g0\ntensor B g3\nncon C A -1,1 B 1,-2,-3\nsave C out result\n".replace("g3", "g6")
    items2 = dict(items)
    b.save_tensor_data("g6", rand_tensor(rng, (2, 3, 2), np.complex128))
    prog = b.compile_program(text)
    prog.run()
    new_a, new_b = rand_tensor(rng, (2, 2), np.complex128), rand_tensor(rng, (2, 3, 2), np.complex128)
    b.save_tensors([("g0", new_a), ("g6", new_b)])
    prog.run()
    ref = layer1.contract_tensors((new_a, new_b), ([-1, 1], [1, -2, -3]))
    assert rel_l2(b.load_tensor_data("result"), ref) < TOL[np.dtype(dtype)]
    # a size change rebinds the label: the program must refuse to run on stale shapes
    from picoquant_jl_b200.host.b200_backend import B200Error
    b.save_tensors([("g0", rand_tensor(rng, (3, 3), np.complex128))])
    with pytest.raises(B200Error):
        prog.run()
    assert items2["g0"].shape == (2, 2)
    # bad arguments leave an error, not a crash
    with pytest.raises(B200Error):
        b.save_tensors([("bad", np.zeros((0, 2)))])


def test_program_view_starts_are_bounds_checked():
    """Caller-supplied per-slice view starts outside [1, ext - nsel + 1] must be rejected on the
    host (BoundsError in the reference, src/layer1.jl:191-194), for run and run_slices."""
    from picoquant_jl_b200.host.b200_backend import B200Error
    b = B200(np.complex128)
    rng = np.random.default_rng(3)
    b.save_tensor_data("t", rand_tensor(rng, (2, 4, 2), np.complex128))
    prog = b.compile_program("tensor A t\nview V A 2 2\nsave V out result\n")
    prog.run([3])
    assert b.load_tensor_data("result").shape == (2, 1, 2)
    for bad in ([0], [5], [-1]):
        with pytest.raises(B200Error):
            prog.run(bad)
    with pytest.raises(B200Error):
        prog.run_slices([[1], [9]], "acc", 2)
    prog.run_slices([[1], [4]], "acc", 2)
    assert b.load_tensor_data("acc").shape == (2, 1, 2)
