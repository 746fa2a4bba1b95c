"""Parity tests of the INT8 tensor-core complex GEMM (csrc/kernels_zgemm_ozaki2.cu, k_ozaki_t)
FORCED onto every eligible step (options ``zgemm_ozaki = 6`` / ``cgemm_ozaki = 4``), i.e. also
onto shapes the default policy leaves to other kernels: small M (one or two tiles, ragged), narrow
N, short and odd K, contracted axis fastest, scattered bits.  The default policy itself is covered
by tests/test_gpu_parity.py::test_default_policy_int8_kernel; the kernel's arithmetic runs on the
CPU in csrc/test_lower (tests/test_abi.py::test_lowering_host_emulation) and its mbarrier protocol
in tests/test_ozaki_t_protocol.py.

Tolerance: 1e-11 rel-L2 per ComplexF64 contraction (north star: 1e-10 on whole flows), 2e-6 per
ComplexF32 contraction against the ComplexF64 oracle (north star: 1e-5)."""
import numpy as np
import pytest

from helpers import rel_l2
from oracle.interactive import OracleBackend
from picoquant_jl_b200.host import create_RQC
from picoquant_jl_b200.host.planner import sweep_plan
from picoquant_jl_b200.host.sliced import SlicedContraction, record_sliced_contraction

pytestmark = pytest.mark.gpu


def B200(dtype=np.complex128, **opts):
    from picoquant_jl_b200.host.b200_backend import B200Backend
    b = B200Backend(dtype)
    for k, v in opts.items():
        b.set_option(k, v)
    return b


SHAPES = [
    # (A dims, A idx, B dims, B idx): canonical, ragged, many tiles, k-fastest, scattered bits
    ((128, 64), [-1, 1], (64, 64), [-2, 1]),
    ((100, 40), [-1, 1], (33, 40), [-2, 1]),
    ((40000, 8), [-1, 1], (17, 8), [-2, 1]),
    ((64, 300), [1, -1], (64, 24), [1, -2]),
    ((5, 1), [-1, 1], (3, 1), [-2, 1]),
    ((2,) * 16, [-1, 1, -2, -3, 2, -4, -5, -6, -7, 3, -8, -9, -10, 4, 5, -11], (2,) * 10,
     [5, 4, 3, 2, 1, -12, -13, -14, -15, -16]),
]


@pytest.mark.parametrize("dtype,opt,groups,tol", [(np.complex128, "zgemm_ozaki", 6, 1e-11),
                                                 (np.complex64, "cgemm_ozaki", 4, 2e-6)])
@pytest.mark.parametrize("case", range(len(SHAPES)))
def test_forced_int8_contraction_matches_oracle(case, dtype, opt, groups, tol):
    ad, ai, bd, bi = SHAPES[case]
    rng = np.random.default_rng(case)
    A = np.asarray((rng.standard_normal(ad) + 1j * rng.standard_normal(ad)).astype(dtype), order="F")
    B = np.asarray((rng.standard_normal(bd) + 1j * rng.standard_normal(bd)).astype(dtype), order="F")
    A[..., 0] *= 1e-3        # unequal row magnitudes
    ref = OracleBackend(np.complex128)
    ref.save_tensor_data("A", A.astype(np.complex128))
    ref.save_tensor_data("B", B.astype(np.complex128))
    ref.contract_tensors("A", ai, "B", bi, "C")
    b = B200(dtype, **{opt: groups})
    b.set_option("fused", 0)
    b.save_tensor_data("A", A)
    b.save_tensor_data("B", B)
    b.profile_enable(True)
    b.contract_tensors("A", ai, "B", bi, "C")
    prof = b.profile_read()
    got = b.load_tensor_data("C")
    assert rel_l2(got, ref.load_tensor_data("C")) < tol, (ad, prof)
    assert b.microbench("ozaki_t_debug") == 0
    b.close()


@pytest.mark.parametrize("case", range(len(SHAPES)))
def test_int8_w_planes_in_tensor_memory(case):
    """Option ``ozaki_tsw = 2``: digit planes 0..3 of W live in tensor memory and their MMAs use the
    TS form of tcgen05.mma (A operand from TMEM); no accumulator is double-buffered then.  The
    integer arithmetic is the same, so the result must be IDENTICAL to the default kernel's."""
    ad, ai, bd, bi = SHAPES[case]
    rng = np.random.default_rng(100 + case)
    A = np.asarray(rng.standard_normal(ad) + 1j * rng.standard_normal(ad), order="F")
    B = np.asarray(rng.standard_normal(bd) + 1j * rng.standard_normal(bd), order="F")
    out = []
    for tsw in (0, 2):
        b = B200(np.complex128, zgemm_ozaki=6, fused=0, ozaki_tsw=tsw)
        for rep in range(2):      # the second launch re-uses tensor memory a first one has written
            b.save_tensor_data("A", A)
            b.save_tensor_data("B", B)
            b.profile_enable(True)
            b.contract_tensors("A", ai, "B", bi, "C")
            prof = b.profile_read()
        if case == 0:
            assert "gemm_int8" in prof, prof
        out.append(np.asarray(b.load_tensor_data("C")).copy())
        assert b.microbench("ozaki_t_debug") == 0
        b.close()
    assert np.array_equal(out[0], out[1])


def test_forced_int8_rejects_other_group_counts():
    from picoquant_jl_b200.host.b200_backend import B200Error
    b = B200()
    with pytest.raises(B200Error):
        b.set_option("zgemm_ozaki", 7)
    with pytest.raises(B200Error):
        b.set_option("cgemm_ozaki", 3)
    b.close()


def test_int8_sliced_rqc_amplitude():
    """A sliced 4x4 depth-12 RQC amplitude: DMMA only (ozaki_auto = 0), the default policy, and
    every eligible GEMM step forced onto the INT8 kernel agree to the north-star tolerance."""
    circ = create_RQC(4, 4, 12, seed=1)
    rec = record_sliced_contraction(circ, 4, 1, plan_fn=lambda tn, s: sweep_plan(tn, 4, 4, sliced_bonds=s),
                                    output_config="0" * 16)
    amps = {}
    for name, opts in (("dmma", dict(ozaki_auto=0)), ("default", dict()), ("forced", dict(zgemm_ozaki=6))):
        b = B200(**opts)
        sc = SlicedContraction(b, rec)
        sc.run([1, 2, 3, 4], "amp")
        amps[name] = complex(np.asarray(b.load_tensor_data("amp")).ravel()[0])
        b.close()
    for name in ("default", "forced"):
        assert abs(amps[name] - amps["dmma"]) / abs(amps["dmma"]) < 1e-10
