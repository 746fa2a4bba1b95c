"""Parity tests of the EXPERIMENTAL INT8 tensor-core ZGEMM (csrc/kernels_zgemm_ozaki.cu,
option ``zgemm_ozaki``).  The kernel has not been brought up on hardware yet
(tools/ozaki_probe.py does that in stages, under timeouts), so these tests are skipped
unless ``PQ_TEST_OZAKI=1`` is set; its arithmetic is covered on the CPU by
csrc/test_lower (tests/test_abi.py::test_lowering_host_emulation).

Tolerance: the north-star's 1e-10 rel-L2 for ComplexF64 on whole flows; 1e-11 per contraction."""
import os

import numpy as np
import pytest

from helpers import rel_l2
from oracle.interactive import OracleBackend
from picoquant_jl_b200.host import create_RQC
from picoquant_jl_b200.host.planner import sweep_plan
from picoquant_jl_b200.host.sliced import SlicedContraction, record_sliced_contraction

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("PQ_TEST_OZAKI") != "1",
                                 reason="experimental kernel: set PQ_TEST_OZAKI=1 after bring-up")]


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
    ((2,) * 16, [-1, 1, -2, -3, 2, -4, -5, -6, -7, 3, -8, -9, -10, 4, 5, -11], (2,) * 10,
     [5, 4, 3, 2, 1, -12, -13, -14, -15, -16]),
]


@pytest.mark.parametrize("groups", [6, 7])
@pytest.mark.parametrize("case", range(len(SHAPES)))
def test_ozaki_contraction_matches_oracle(case, groups):
    ad, ai, bd, bi = SHAPES[case]
    rng = np.random.default_rng(case)
    A = np.asarray(rng.standard_normal(ad) + 1j * rng.standard_normal(ad), order="F")
    B = np.asarray(rng.standard_normal(bd) + 1j * rng.standard_normal(bd), order="F")
    A[..., 0] *= 1e-3        # unequal row magnitudes
    ref = OracleBackend(np.complex128)
    ref.save_tensor_data("A", A)
    ref.save_tensor_data("B", B)
    ref.contract_tensors("A", ai, "B", bi, "C")
    b = B200(zgemm_ozaki=groups)
    b.save_tensor_data("A", A)
    b.save_tensor_data("B", B)
    b.contract_tensors("A", ai, "B", bi, "C")
    assert rel_l2(b.load_tensor_data("C"), ref.load_tensor_data("C")) < 1e-11
    b.close()


@pytest.mark.parametrize("case", range(len(SHAPES)))
def test_ozaki_c64_contraction_matches_f64_oracle(case):
    """ComplexF32 twin (option cgemm_ozaki = 4): within 1e-5 of the ComplexF64 oracle downcast,
    and no worse than 2e-6 on these well-scaled operands."""
    ad, ai, bd, bi = SHAPES[case]
    rng = np.random.default_rng(100 + case)
    A = np.asarray((rng.standard_normal(ad) + 1j * rng.standard_normal(ad)).astype(np.complex64), order="F")
    B = np.asarray((rng.standard_normal(bd) + 1j * rng.standard_normal(bd)).astype(np.complex64), order="F")
    ref = OracleBackend(np.complex128)
    ref.save_tensor_data("A", A.astype(np.complex128))
    ref.save_tensor_data("B", B.astype(np.complex128))
    ref.contract_tensors("A", ai, "B", bi, "C")
    b = B200(np.complex64, cgemm_ozaki=4)
    b.save_tensor_data("A", A)
    b.save_tensor_data("B", B)
    b.contract_tensors("A", ai, "B", bi, "C")
    assert rel_l2(b.load_tensor_data("C"), ref.load_tensor_data("C")) < 2e-6
    b.close()


LONG_K = [
    # canonical TTGT + k_zgemm_ozaki_kloop: 64 < K <= 8192
    ((300, 200), [-1, 1], (70, 200), [-2, 1]),
    ((64, 1100, 8), [-1, 1, -2], (1100, 40), [1, -3]),
    ((2,) * 19, [1, -1, 2, -2, 3, -3, 4, -4, 5, -5, 6, -6, 7, -7, 8, -8, 9, 10, 11], (2,) * 16,
     [11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, -9, -10, -11, -12, -13]),
]


@pytest.mark.parametrize("dtype,groups,tol", [(np.complex128, 6, 1e-11), (np.complex128, 7, 1e-11),
                                             (np.complex64, 4, 2e-6)])
@pytest.mark.parametrize("case", range(len(LONG_K)))
def test_ozaki_long_contraction_matches_oracle(case, dtype, groups, tol):
    ad, ai, bd, bi = LONG_K[case]
    rng = np.random.default_rng(200 + case)
    A = np.asarray((rng.standard_normal(ad) + 1j * rng.standard_normal(ad)).astype(dtype), order="F")
    B = np.asarray((rng.standard_normal(bd) + 1j * rng.standard_normal(bd)).astype(dtype), order="F")
    ref = OracleBackend(np.complex128)
    ref.save_tensor_data("A", A.astype(np.complex128))
    ref.save_tensor_data("B", B.astype(np.complex128))
    ref.contract_tensors("A", ai, "B", bi, "C")
    opt = "zgemm_ozaki" if dtype == np.complex128 else "cgemm_ozaki"
    b = B200(dtype, **{opt: groups})
    b.save_tensor_data("A", A)
    b.save_tensor_data("B", B)
    b.profile_enable(True)
    b.contract_tensors("A", ai, "B", bi, "C")
    assert "gemm_int8" in b.profile_read()
    assert rel_l2(b.load_tensor_data("C"), ref.load_tensor_data("C")) < tol
    b.close()


def test_ozaki_sliced_rqc_amplitude():
    """A sliced 4x4 depth-12 RQC amplitude with every eligible GEMM step on the INT8 kernel."""
    circ = create_RQC(4, 4, 12, seed=1)
    rec = record_sliced_contraction(circ, 4, 1, plan_fn=lambda tn, s: sweep_plan(tn, 4, 4, sliced_bonds=s),
                                    output_config="0" * 16)
    amps = {}
    for g in (0, 6, 7):
        b = B200(zgemm_ozaki=g)
        sc = SlicedContraction(b, rec)
        sc.run([1, 2, 3, 4], "amp")
        amps[g] = complex(np.asarray(b.load_tensor_data("amp")).ravel()[0])
        b.close()
    for g in (6, 7):
        assert abs(amps[g] - amps[0]) / abs(amps[0]) < 1e-10
