"""Pins the CPU oracle (oracle/) and the host mirror against every golden /
known-answer the reference holds for this path.  CPU only."""
import json
import random

import numpy as np
import pytest

from helpers import (GOLDEN, golden_qasm, load_golden_json, rel_l2, statevector,
                     switch_endianness)
from oracle import layer1
from oracle.interactive import OracleBackend, execute_dsl
from picoquant_jl_b200.host import (Circuit, DSLBackend, TensorNetworkCircuit, add_gate,
                                    add_input, add_output, contract_network, contract_pair,
                                    convert_circuit_to_network, create_ghz_preparation_circuit,
                                    create_qft_circuit, create_RQC,
                                    create_simple_preparation_circuit,
                                    full_wavefunction_contraction, gate_tensor,
                                    load_qasm_as_circuit, network_from_dict,
                                    partition_network_on_virtual_bonds,
                                    random_contraction_plan, slice_tensor_network)

C128 = np.complex128

def _asf(a):
    """Fortran-contiguous view/copy that keeps 0-d arrays 0-d."""
    return np.asarray(a, order="F")



def test_reference_stored_contraction_golden():
    """examples/ghz_3.json + ghz_3_plan.json -> ghz_3_contracted.json: the only
    place the reference stores the *values and layout* of a contracted tensor.
    Index order [A-open..., B-open...] and column-major data must match exactly."""
    d = load_golden_json("ghz_3.json")
    g = load_golden_json("ghz_3_contracted.json")
    b = OracleBackend(C128)
    tn = network_from_dict(d, b)
    for k, v in d["nodes"].items():
        data = np.array(v["data_re"]) + 1j * np.array(v["data_im"])
        b.save_tensor_data(k, np.reshape(data, v["data_dims"], order="F"))
    for edge in load_golden_json("ghz_3_plan.json"):
        contract_pair(tn, edge)
    (label, gnode), = g["nodes"].items()
    assert list(tn.nodes) == [label]
    assert tn.nodes[label].indices == gnode["indices"]
    out = b.load_tensor_data(label)
    assert list(out.shape) == gnode["data_dims"]
    ref = np.array(gnode["data_re"]) + 1j * np.array(gnode["data_im"])
    assert np.array_equal(out.ravel(order="F"), ref)
    for k, e in g["edges"].items():
        assert (tn.edges[k].src, tn.edges[k].dst) == (e["src"], e["dst"])


def test_metrics_golden_8_44_124():
    """test/layer2_tests.jl:106-143 (h, cx, cx, cx(0,2); full-wf)."""
    qasm = """OPENQASM 2.0;
              include "qelib1.inc";
              qreg q[3];
              h q[0];
              cx q[0],q[1];
              cx q[0],q[1];
              cx q[0],q[2];"""
    for backend in (OracleBackend(), DSLBackend()):
        tn = convert_circuit_to_network(load_qasm_as_circuit(qasm), backend)
        add_input(tn, "000")
        full_wavefunction_contraction(tn, "vector")
        assert backend.metrics.as_tuple() == (8, 44, 124)
        assert len(tn.nodes) == 1


def test_ghz3_amplitude_random_plan_through_dsl():
    """test/layer1_tests.jl:11-46: <000|GHZ3> = 1/sqrt(2) via a random edge
    plan, DSL stream, then the DSL interpreter in ComplexF64."""
    circ = load_qasm_as_circuit(golden_qasm("ghz_3.qasm"))
    for seed in range(5):
        dsl = DSLBackend()
        tn = convert_circuit_to_network(circ, dsl)
        add_input(tn, "000")
        add_output(tn, "000")
        contract_network(tn, random_contraction_plan(tn, random.Random(seed)))
        assert len(tn.nodes) == 1
        execute_dsl(dsl.text(), dsl.store, C128)
        res = dsl.load_tensor_data("result")
        assert res.shape == ()
        assert abs(res - 1 / np.sqrt(2)) < 1e-12


def test_disjoint_network_vector_result():
    """test/layer2_tests.jl:69-103: H⊗H|00>, leftover pieces are contracted in
    insertion order, result is 1-D with first amplitude 1/2."""
    circ = Circuit(2).h(0).h(1)
    for seed in range(3):
        b = OracleBackend()
        tn = convert_circuit_to_network(circ, b)
        add_input(tn, "00")
        contract_network(tn, random_contraction_plan(tn, random.Random(seed)), "vector")
        assert len(tn.nodes) == 1
        res = b.load_tensor_data("result")
        assert res.ndim == 1
        assert abs(res.real[0] - 0.5) < 1e-6


def test_ghz5_state():
    """test/layer2_tests.jl:308-328 known answer (via full-wf here; the MPS path
    is out of scope): 1/sqrt(2) at the first and last amplitude."""
    b = OracleBackend(C128)
    psi = statevector(create_ghz_preparation_circuit(5), b)
    ref = np.zeros(32, dtype=C128)
    ref[[0, -1]] = 1 / np.sqrt(2)
    assert rel_l2(psi, ref) < 1e-14


def test_decomposed_gate_recontracts():
    """test/layer3_tests.jl:57-71: contracting the two SVD halves gives the gate
    back after permutedims(.., (1,3,2,4)) -- pins C's axis order."""
    rng = np.random.default_rng(7)
    gate = _asf(rng.standard_normal((2, 2, 2, 2)) + 1j * rng.standard_normal((2, 2, 2, 2)))
    b = OracleBackend(C128)
    tn = TensorNetworkCircuit(2, b)
    labels = add_gate(tn, gate, [1, 2], decompose=True)
    assert len(labels) == 2
    out = contract_pair(tn, *labels)
    data = np.transpose(b.load_tensor_data(out), (0, 2, 1, 3))
    assert rel_l2(data, gate) < 1e-13


def test_decompose_counts():
    """test/layer3_tests.jl:73-87 and :38-55."""
    circ = load_qasm_as_circuit(golden_qasm("ghz_3.qasm"))
    tn = convert_circuit_to_network(circ, OracleBackend(), decompose=True)
    assert len(tn.nodes) == 5
    tn = TensorNetworkCircuit(3, OracleBackend())
    assert len(tn.nodes) == 0
    add_gate(tn, gate_tensor("H"), [1])
    assert len(tn.nodes) == 1 and len(tn.edges) == 4


def test_add_input_output_idempotent():
    """test/layer3_tests.jl:105-124."""
    tn = TensorNetworkCircuit(3, OracleBackend())
    add_input(tn, "000")
    assert len(tn.nodes) == 3
    add_output(tn, "000")
    assert len(tn.nodes) == 6
    add_input(tn, "000")
    add_output(tn, "000")
    assert len(tn.nodes) == 6


@pytest.mark.parametrize("n", [3, 8])
def test_qft_against_inverse_fft(n):
    """test/algorithms_tests.jl:39-82: prep circuit followed by QFT equals the
    normalised inverse FFT of the prep state (both big-endian)."""
    prep = create_simple_preparation_circuit(n, 3, 43)
    full = prep.compose(create_qft_circuit(n))
    psi_in = statevector(prep, OracleBackend(C128))
    assert rel_l2(psi_in, prep.to_matrix()[:, 0]) < 1e-13   # stands in for qiskit Aer
    ref = np.fft.ifft(switch_endianness(psi_in))
    ref /= np.linalg.norm(ref)
    psi = switch_endianness(statevector(full, OracleBackend(C128)))
    assert abs(abs(np.vdot(psi, ref)) - 1.0) < 1e-12


@pytest.mark.parametrize("iswap", [False, True])
def test_small_rqc_against_dense_simulation(iswap):
    """test/algorithms_tests.jl:84-115 (qiskit Aer replaced by a dense unitary)."""
    rqc = create_RQC(3, 3, 8, seed=11, use_iswap=iswap, final_Hadamard_layer=iswap)
    psi = statevector(rqc, OracleBackend(C128))
    assert rel_l2(psi, rqc.to_matrix()[:, 0]) < 1e-12


@pytest.mark.parametrize("name,n", [("qft_2.qasm", 2), ("qft_3.qasm", 3), ("qft_5.qasm", 5),
                                    ("qft_10.qasm", 10)])
def test_qft_fixtures_uniform_superposition(name, n):
    """QFT|0..0> is the uniform superposition (closed form, SURVEY §8d cfg 2);
    with a '1' input the fixture must agree with the dense unitary."""
    circ = load_qasm_as_circuit(golden_qasm(name))
    psi = statevector(circ, OracleBackend(C128))
    assert rel_l2(psi, np.full(2 ** n, 2 ** (-n / 2))) < 1e-12
    if n <= 5:
        cfg = "1" + "0" * (n - 1)
        psi1 = statevector(circ, OracleBackend(C128), input_config=cfg)
        assert rel_l2(psi1, circ.to_matrix()[:, 1]) < 1e-12


def test_slicing_identity_sum_of_slices():
    """test/layer2_tests.jl:419-455: the sum over P=4 slices equals the unsliced
    wavefunction (transpile dropped: needs qiskit)."""
    n = 4
    circ = create_simple_preparation_circuit(n, 2, 5).compose(create_qft_circuit(n))
    wf = statevector(circ, OracleBackend(C128), decompose=True)
    for P in (2, 4, 8):
        total = np.zeros_like(wf)
        for p in range(1, P + 1):
            b = OracleBackend(C128)
            tn = convert_circuit_to_network(circ, b, decompose=True)
            add_input(tn, "0" * n)
            labels, values = partition_network_on_virtual_bonds(tn, P, p)
            assert len(labels) == int(np.log2(P))
            slice_tensor_network(tn, labels, values)
            full_wavefunction_contraction(tn, "vector")
            total += b.load_tensor_data("result")
        assert rel_l2(total, wf) < 1e-13


def test_oracle_layer1_units():
    """test/layer1_tests.jl:48-64."""
    A = _asf(np.array([[1 + 1j, 1j], [-1j, 2.0]]))
    t = layer1.conjugate_tensor(layer1.transpose_tensor(A, [2, 1]))
    assert np.array_equal(t, A.conj().T)
    v = layer1.reshape_tensor(t, 4)
    assert np.array_equal(v, np.array([1 - 1j, -1j, 1j, 2.0]))


def test_oracle_contract_matches_einsum():
    rng = np.random.default_rng(0)
    A = _asf(rng.standard_normal((2, 3, 4, 2)) + 1j * rng.standard_normal((2, 3, 4, 2)))
    B = _asf(rng.standard_normal((4, 5, 3)) + 1j * rng.standard_normal((4, 5, 3)))
    C = layer1.contract_tensors((A, B), ([-1, 2, 1, -2], [1, -3, 2]))
    ref = np.einsum("abcd,ceb->ade", A, B)
    assert C.flags.f_contiguous and C.shape == (2, 2, 5)
    assert rel_l2(C, ref) < 1e-14
    # view keeps the axis, 1-based
    V = layer1.tensor_view(A, 3, range(2, 3))
    assert V.shape == (2, 3, 1, 2) and np.array_equal(V[:, :, 0, :], A[:, :, 1, :])


def test_qasm_load_and_json_round_trip():
    """test/layer3_tests.jl:10-36: qasm -> circuit -> network; to_json(network_from_json(j)) == j."""
    from picoquant_jl_b200.host import network_from_json, to_json
    qasm = """OPENQASM 2.0;
              include "qelib1.inc";
              qreg q[3];
              h q[0];
              cx q[0],q[1];
              cx q[1],q[2];"""
    circ = load_qasm_as_circuit(qasm)
    assert circ.n_qubits == 3 and len(circ.data) == 3
    tng = convert_circuit_to_network(circ, OracleBackend(C128))
    j = to_json(tng)
    assert to_json(network_from_json(j, OracleBackend(C128))) == j


def test_network_data_structure_counts():
    """test/layer3_tests.jl:38-55: empty network, one node and four edges after one
    single-qubit gate on a 3-qubit register."""
    tn = TensorNetworkCircuit(3, OracleBackend(C128))
    assert len(tn.nodes) == 0
    add_gate(tn, np.array([[1, 1], [1, -1]]) / np.sqrt(2), [1])
    assert len(tn.nodes) == 1
    assert len(tn.edges) == 4


def test_transpile_qubit_ordering_golden_and_state():
    """test/layer3_tests.jl:89-102: ``h q0; cx q0,q2`` transpiled onto a line gives
    qubit_ordering == [2, 1, 3].  And routing must not change the state: QFT-4/5 after a
    preparation layer, plain and through the MPS driver (which needs neighbouring gates)."""
    from picoquant_jl_b200.host import (calculate_mps_amplitudes, contract_mps_tensor_network_circuit,
                                        transpile_circuit)
    qasm = """OPENQASM 2.0;
              include "qelib1.inc";
              qreg q[3];
              h q[0];
              cx q[0],q[2];"""
    tng = convert_circuit_to_network(load_qasm_as_circuit(qasm), OracleBackend(C128), transpile=True)
    assert tng.qubit_ordering == [2, 1, 3]
    routed, order = transpile_circuit(load_qasm_as_circuit(qasm))
    assert [g[0] for g in routed.gates()] == ["h", "swap", "cx"] and order == [2, 1, 3]
    assert routed.gates()[1][2] == (0, 1) and routed.gates()[2][2] == (1, 2)
    for n in (4, 5):
        circ = create_simple_preparation_circuit(n, 2, 3).compose(create_qft_circuit(n))
        ref = statevector(circ, OracleBackend(C128))
        assert rel_l2(statevector(circ, OracleBackend(C128), transpile=True), ref) < 1e-12
        routed, _ = transpile_circuit(circ)
        assert all(abs(q[0] - q[1]) == 1 for _, _, q in routed.gates() if len(q) == 2)
        b = OracleBackend(C128)
        tn = convert_circuit_to_network(circ, b, decompose=True, transpile=True)
        add_input(tn, "0" * n)
        mps_nodes = contract_mps_tensor_network_circuit(tn)
        calculate_mps_amplitudes(tn, mps_nodes)
        assert rel_l2(b.load_tensor_data("result"), ref) < 1e-10
