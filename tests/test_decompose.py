"""decompose_tensor! / compress_bond! / compress_tensor_chain!.

CPU part: the oracle restatement of ``src/layer1.jl:146-184`` against the reference's own
tests for this function (``test/layer2_tests.jl:241-305`` compression of a tensor chain,
``:355-418`` threshold and max_rank) and against closed-form properties of the SVD split.
GPU part: the device implementation (``pq_decompose``, one-sided Jacobi SVD) against the
oracle on the same inputs.  U and V of an SVD are unique only up to phases (and rotations
inside degenerate subspaces), so parity is checked on what IS unique: chi, the singular
values (B^H B = C C^H = diag(S)), and the product B*C = best rank-chi approximation.
"""
import numpy as np
import pytest

from helpers import TOL, rel_l2
from oracle import layer1
from oracle.interactive import OracleBackend
from picoquant_jl_b200.host import (DSLBackend, TensorNetworkCircuit, add_gate, add_input,
                                    calculate_mps_amplitudes, compress_tensor_chain,
                                    contract_mps_tensor_network_circuit, contract_pair,
                                    convert_circuit_to_network, create_ghz_preparation_circuit,
                                    create_qft_circuit, create_simple_preparation_circuit,
                                    decompose_tensor, full_wavefunction_contraction,
                                    inorder_contraction, load_qasm_as_circuit, virtualedges)

DTYPES = [np.complex128, np.complex64]


def rand_tensor(rng, shape, dtype):
    a = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    return np.asarray(a.astype(dtype), order="F")


def haar_unitary(rng, n):
    z = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))) / np.sqrt(2)
    q, r = np.linalg.qr(z)
    return q * (np.diag(r) / np.abs(np.diag(r)))


def as_matrix(t, left, right):
    dims = t.shape
    p = np.transpose(t, [x - 1 for x in list(left) + list(right)])
    m = int(np.prod([dims[x - 1] for x in left], dtype=np.int64))
    return np.reshape(np.asarray(p, order="F"), (m, -1), order="F")


def factors_as_matrices(B, C):
    chi = B.shape[-1]
    return (np.reshape(np.asarray(B, order="F"), (-1, chi), order="F"),
            np.reshape(np.asarray(C, order="F"), (chi, -1), order="F"))


def check_split(t, left, right, B, C, chi, threshold, max_rank, tol):
    """Properties every correct decompose_tensor result has (layer1.jl:146-184)."""
    A = as_matrix(t.astype(np.complex128), left, right)
    S = np.linalg.svd(A, compute_uv=False)
    eps = np.finfo(t.dtype).eps
    thr = max(threshold, float(np.sqrt(eps)))
    expect = int(np.sum(S / np.sqrt(np.sum(S ** 2)) > thr))
    if max_rank > 0:
        expect = min(expect, max_rank)
    assert chi == expect, (chi, expect, S)
    assert B.shape == tuple(t.shape[x - 1] for x in left) + (chi,)
    assert C.shape == (chi,) + tuple(t.shape[x - 1] for x in right)
    Bm, Cm = factors_as_matrices(B.astype(np.complex128), C.astype(np.complex128))
    scale = S[0] if len(S) and S[0] > 0 else 1.0
    # both factors carry sqrt(S): B^H B = C C^H = diag(S[:chi])
    assert np.linalg.norm(Bm.conj().T @ Bm - np.diag(S[:chi])) <= 20 * tol * scale
    assert np.linalg.norm(Cm @ Cm.conj().T - np.diag(S[:chi])) <= 20 * tol * scale
    # B C is the rank-chi truncation of A: the error is exactly the discarded tail
    tail = np.sqrt(np.sum(S[chi:] ** 2))
    err = np.linalg.norm(Bm @ Cm - A)
    assert abs(err - tail) <= 20 * tol * np.linalg.norm(A) + 1e-300, (err, tail)


CASES = [
    # shape, left positions, right positions, threshold, max_rank
    ((2, 2, 2, 2), [1, 3], [2, 4], 1e-13, 0),          # a two-qubit gate (decompose_gate! shape)
    ((2, 2, 2, 2), [1, 3], [2, 4], 0.5, 0),            # layer2_tests.jl:355-386
    ((2, 2, 2, 2), [1, 3], [2, 4], 1e-13, 1),          # layer2_tests.jl:388-418
    ((4, 3, 5), [2], [3, 1], 1e-13, 0),                # m < n, odd sizes
    ((6, 7, 2), [1, 2], [3], 1e-13, 0),                # m > n
    ((5, 1, 3), [1], [2, 3], 1e-13, 0),                # extent-1 axis
    ((8, 4, 8, 4), [1, 2], [3, 4], 1e-13, 5),          # MPS bond, max_rank below full rank
    ((2, 16, 2, 16), [2, 1], [3, 4], 0.05, 0),         # relative threshold cuts the tail
    ((7,), [1], [], 1e-13, 0),                          # empty right side: n = 1
    ((40, 90), [1], [2], 1e-13, 0),                    # many columns: one launch per Jacobi step
    ((150, 70), [2], [1], 1e-13, 33),
]


@pytest.mark.parametrize("dtype", DTYPES)
def test_oracle_decompose_properties(dtype):
    rng = np.random.default_rng(5)
    for shape, left, right, thr, mr in CASES:
        t = rand_tensor(rng, shape, dtype)
        B, C, chi = layer1.decompose_tensor(t, left, right, thr, mr)
        assert B.dtype == np.dtype(dtype) and C.dtype == np.dtype(dtype)
        check_split(t, left, right, B, C, chi, thr, mr, TOL[np.dtype(dtype)])


def test_oracle_decompose_rank_deficient_and_zero():
    rng = np.random.default_rng(6)
    u = rand_tensor(rng, (12, 3), np.complex128)
    v = rand_tensor(rng, (3, 10), np.complex128)
    t = np.asarray((u @ v).reshape((12, 2, 5), order="F"), order="F")
    B, C, chi = layer1.decompose_tensor(t, [1], [2, 3])
    assert chi == 3
    check_split(t, [1], [2, 3], B, C, chi, 1e-13, 0, 1e-10)
    z = np.zeros((3, 4), dtype=np.complex128, order="F")
    B, C, chi = layer1.decompose_tensor(z, [1], [2])
    assert chi == 0 and B.shape == (3, 0) and C.shape == (0, 4)   # NaN > thr is false in Julia too


def _threshold_and_max_rank(backend_factory):
    """test/layer2_tests.jl:355-418 -- Haar-random two-qubit gate."""
    rng = np.random.default_rng(7)
    d = 2
    gate = np.reshape(haar_unitary(rng, d * d), (d, d, d, d), order="F")
    F = np.linalg.svd(np.reshape(np.transpose(gate, (0, 2, 1, 3)), (d * d, d * d), order="F"),
                      compute_uv=False)
    for kw, expected in ((dict(threshold=0.5), int(np.sum(F / np.sqrt(np.sum(F ** 2)) > 0.5))),
                         (dict(max_rank=1), 1)):
        b = backend_factory()
        tn = TensorNetworkCircuit(2, b)
        add_gate(tn, gate, [1, 2])
        new_nodes = decompose_tensor(tn, "node_1", ["index_1", "index_3"], ["index_2", "index_4"], **kw)
        bond = virtualedges(tn, new_nodes[0])[0]
        idx = tn.nodes[new_nodes[0]].indices.index(bond)
        assert b.load_tensor_data(new_nodes[0]).shape[idx] == expected
        assert tn.nodes[new_nodes[0]].dims[idx] == expected
        assert b.load_tensor_data("node_1") is None          # the original is consumed
        e = tn.edges[bond]
        assert (e.src, e.dst, e.virtual) == (new_nodes[0], new_nodes[1], True)


def _compress_chain(backend_factory):
    """test/layer2_tests.jl:255-305: h, cx, cx, cx on two qubits; each gate is split, all
    non-virtual closed edges are contracted, leaving two nodes joined by three virtual bonds
    of dimension 2; compressing the chain must leave bond dimension 2."""
    qasm = """OPENQASM 2.0;
              include "qelib1.inc";
              qreg q[2];
              h q[0];
              cx q[0],q[1];
              cx q[0],q[1];
              cx q[0],q[1];
              """
    b = backend_factory()
    tn = convert_circuit_to_network(load_qasm_as_circuit(qasm), b)
    add_input(tn, "00")
    decompose_tensor(tn, "node_2", ["index_3", "index_4"], ["index_2", "index_5"])
    decompose_tensor(tn, "node_3", ["index_4", "index_6"], ["index_5", "index_7"])
    decompose_tensor(tn, "node_4", ["index_6", "index_8"], ["index_7", "index_9"])
    plan = [k for k, v in tn.edges.items() if not v.virtual and v.src is not None and v.dst is not None]
    for edge in plan:
        contract_pair(tn, edge)
    assert list(tn.nodes) == ["node_18", "node_19"]
    before = np.array(_state_of_two_nodes(tn, b))
    compress_tensor_chain(tn, ["node_18", "node_19"])
    idx = tn.nodes["node_18"].indices.index("index_14")
    result = b.load_tensor_data("node_18")
    assert result.shape[idx] == 2
    # the state itself is unchanged by the compression: (|00> + |11>)/sqrt(2)
    after = np.array(_state_of_two_nodes(tn, b))
    assert rel_l2(after, before) < 1e-5
    assert abs(abs(after[0, 0]) - 1 / np.sqrt(2)) < 1e-5 and abs(abs(after[1, 1]) - 1 / np.sqrt(2)) < 1e-5


def _state_of_two_nodes(tn, b):
    """Contracts the two remaining nodes on the host (without touching the network)."""
    A, B = (np.asarray(b.load_tensor_data(k)) for k in ("node_18", "node_19"))
    ia, ib = tn.nodes["node_18"].indices, tn.nodes["node_19"].indices
    shared = [x for x in ia if x in ib]
    out = np.tensordot(A, B, axes=([ia.index(x) for x in shared], [ib.index(x) for x in shared]))
    labels = [x for x in ia if x not in shared] + [x for x in ib if x not in shared]
    order = [labels.index(x) for x in tn.output_qubits]
    return np.transpose(out, order)


def _mps_matches_full_wavefunction(backend_factory, tol):
    """test/layer2_tests.jl:204-237 (preparation + GHZ on 3 qubits: MPS == full wave
    function) and :307-326 (GHZ-5 through the MPS path = (|0..0> + |1..1>)/sqrt(2));
    plus a nearest-neighbour circuit with non-trivial bond growth and a max_rank cut."""
    circ = create_simple_preparation_circuit(3, 1).compose(create_ghz_preparation_circuit(3))
    b = backend_factory()
    tn = convert_circuit_to_network(circ, b, decompose=True)
    add_input(tn, "000")
    full_wavefunction_contraction(tn, "vector")
    full_wf = np.array(b.load_tensor_data("result"))
    b = backend_factory()
    tn = convert_circuit_to_network(circ, b, decompose=True)
    add_input(tn, "000")
    mps_nodes = contract_mps_tensor_network_circuit(tn)
    calculate_mps_amplitudes(tn, mps_nodes)
    mps_wf = np.array(b.load_tensor_data("result"))
    assert mps_wf.shape == full_wf.shape == (8,)
    assert rel_l2(mps_wf, full_wf) < 10 * tol

    n = 5
    b = backend_factory()
    tn = convert_circuit_to_network(create_ghz_preparation_circuit(n), b, decompose=True)
    add_input(tn, "0" * n)
    mps_nodes = contract_mps_tensor_network_circuit(tn)
    for k, node in enumerate(mps_nodes):     # bond dimension 2 everywhere, saved under its label
        shape = b.load_tensor_data(node).shape
        assert sorted(shape) == ([2, 2] if k in (0, n - 1) else [2, 2, 2]), (node, shape)
    calculate_mps_amplitudes(tn, mps_nodes)
    ref = np.zeros(2 ** n, dtype=np.complex128)
    ref[[0, -1]] = 1 / np.sqrt(2)
    assert rel_l2(b.load_tensor_data("result"), ref) < 10 * tol


def _inorder_with_bond_merging(backend_factory, tol):
    """``inorder_contraction!`` + ``merge_common_bonds!`` (src/layer2.jl:25-112; the reference
    has no test for them): on a decomposed 4-qubit preparation + QFT network every qubit's
    world-line collapses into one node, multiple virtual bonds between two nodes are fused by
    permute + reshape on the backend, and contracting what is left reproduces the state."""
    circ = create_simple_preparation_circuit(4, 2, 3).compose(create_qft_circuit(4))
    ob = OracleBackend(np.complex128)
    tn = convert_circuit_to_network(circ, ob, decompose=True)
    add_input(tn, "0000")
    full_wavefunction_contraction(tn, "vector")
    ref = np.array(ob.load_tensor_data("result"))

    b = backend_factory()
    tn = convert_circuit_to_network(circ, b, decompose=True)
    add_input(tn, "0000")
    inorder_contraction(tn)
    labels = list(tn.nodes)
    assert len(labels) == 4
    for i, x in enumerate(labels):
        assert list(b.load_tensor_data(x).shape) == tn.nodes[x].dims      # graph dims == data dims
        for y in labels[i + 1:]:
            assert len(set(tn.nodes[x].indices) & set(tn.nodes[y].indices)) <= 1
    assert sorted(max(v.dims) for v in tn.nodes.values()) == [8, 8, 16, 16]   # fused bonds
    for e in tn.edges.values():
        if e.src is not None and e.dst is not None:
            assert e.virtual
    sites = [tn.edges[x].src for x in tn.output_qubits]
    calculate_mps_amplitudes(tn, sites)
    assert rel_l2(b.load_tensor_data("result"), ref) < 10 * tol


def test_inorder_contraction_with_bond_merging_oracle():
    _inorder_with_bond_merging(lambda: OracleBackend(np.complex128), 1e-10)


def test_mps_state_amplitudes_oracle():
    """src/mps.jl + test/layer2_tests.jl:330-351: GHZ-5 amplitudes through the array
    interface of MPSState; and, for a state without the GHZ symmetry, every amplitude equals
    the corresponding entry of calculate_mps_amplitudes!."""
    import itertools
    from picoquant_jl_b200.host import MPSState
    n = 5
    b = OracleBackend(np.complex128)
    tn = convert_circuit_to_network(create_ghz_preparation_circuit(n), b, decompose=True)
    add_input(tn, "0" * n)
    mps_nodes = contract_mps_tensor_network_circuit(tn)
    state = MPSState(tn, mps_nodes)
    assert state.shape == (2,) * n and len(state) == 2 ** n
    assert abs(state["11111"] - 1 / np.sqrt(2)) < 1e-6 and abs(state["00000"] - 1 / np.sqrt(2)) < 1e-6
    assert abs(state["10101"]) < 1e-6
    assert abs(state[2, 2, 2, 2, 2] - state["11111"]) == 0

    circ = create_simple_preparation_circuit(3, 1).compose(create_ghz_preparation_circuit(3))
    b = OracleBackend(np.complex128)
    tn = convert_circuit_to_network(circ, b, decompose=True)
    add_input(tn, "000")
    mps_nodes = contract_mps_tensor_network_circuit(tn)
    state = MPSState(tn, mps_nodes, dtype=np.complex128)
    calculate_mps_amplitudes(tn, mps_nodes)
    full = np.reshape(np.array(b.load_tensor_data("result")), (2, 2, 2), order="F")
    for bits in itertools.product([0, 1], repeat=3):
        assert abs(state["".join(str(x) for x in bits)] - full[bits]) < 1e-12, bits


def test_reference_mps_contraction_oracle():
    _mps_matches_full_wavefunction(lambda: OracleBackend(np.complex128), 1e-10)


def test_reference_threshold_and_max_rank_oracle():
    _threshold_and_max_rank(lambda: OracleBackend(np.complex128))


def test_reference_compress_chain_oracle():
    _compress_chain(lambda: OracleBackend(np.complex128))


def test_dsl_stream_of_decompose():
    """dsl.jl:181-195: the command text, and chi = 0 => graph-side upper bound."""
    b = DSLBackend()
    tn = TensorNetworkCircuit(2, b)
    add_gate(tn, np.reshape(np.eye(4), (2, 2, 2, 2)), [1, 2])
    decompose_tensor(tn, "node_1", ["index_1", "index_3"], ["index_2", "index_4"], max_rank=3)
    lines = b.text().splitlines()
    assert lines[-2] == 'decompose node_1 node_2 1,3 node_3 2,4 {"threshold":1.0e-13, "max_rank":3}'
    assert lines[-1] == "del node_1"
    assert tn.nodes["node_2"].dims == [2, 2, 3] and tn.nodes["node_3"].dims == [3, 2, 2]


# ---------------------------------------------------------------------------------------
# device
# ---------------------------------------------------------------------------------------
def _b200(dtype):
    from picoquant_jl_b200.host.b200_backend import B200Backend
    return B200Backend(dtype)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", DTYPES)
def test_gpu_decompose_matches_oracle(dtype):
    rng = np.random.default_rng(5)
    tol = TOL[np.dtype(dtype)]
    b = _b200(dtype)
    for shape, left, right, thr, mr in CASES:
        t = rand_tensor(rng, shape, dtype)
        b.save_tensor_data("T", t)
        chi = b.decompose_tensor("T", left, right, threshold=thr, max_rank=mr,
                                 left_label="L", right_label="R")
        assert b.load_tensor_data("T") is None
        B, C = b.load_tensor_data("L"), b.load_tensor_data("R")
        Bo, Co, chio = layer1.decompose_tensor(t, left, right, thr, mr)
        assert chi == chio, (shape, chi, chio)
        check_split(t, left, right, B, C, chi, thr, mr, tol)
        # same truncated product as the oracle (unique whenever S[chi-1] > S[chi])
        Bm, Cm = factors_as_matrices(B, C)
        Bom, Com = factors_as_matrices(Bo, Co)
        assert rel_l2(Bm @ Cm, Bom @ Com) < 10 * tol, (shape, rel_l2(Bm @ Cm, Bom @ Com))


@pytest.mark.gpu
def test_gpu_decompose_rank_deficient_zero_and_errors():
    from picoquant_jl_b200.host.b200_backend import B200Error
    rng = np.random.default_rng(6)
    b = _b200(np.complex128)
    u = rand_tensor(rng, (12, 3), np.complex128)
    v = rand_tensor(rng, (3, 10), np.complex128)
    t = np.asarray((u @ v).reshape((12, 2, 5), order="F"), order="F")
    b.save_tensor_data("T", t)
    chi = b.decompose_tensor("T", [1], [2, 3], left_label="L", right_label="R")
    assert chi == 3
    check_split(t, [1], [2, 3], b.load_tensor_data("L"), b.load_tensor_data("R"), chi, 1e-13, 0, 1e-10)
    b.save_tensor_data("Z", np.zeros((3, 4), dtype=np.complex128, order="F"))
    assert b.decompose_tensor("Z", [1], [2], left_label="L", right_label="R") == 0
    assert b.load_tensor_data("L").shape == (3, 0) and b.load_tensor_data("R").shape == (0, 4)
    b.save_tensor_data("T", t)
    with pytest.raises(B200Error):
        b.decompose_tensor("T", [1], [2], left_label="L", right_label="R")       # axis 3 missing
    with pytest.raises(B200Error):
        b.decompose_tensor("T", [1, 1], [2], left_label="L", right_label="R")    # repeated axis
    with pytest.raises(KeyError):
        b.decompose_tensor("nope", [1], [2], left_label="L", right_label="R")


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", DTYPES)
def test_gpu_reference_decompose_tests(dtype):
    _threshold_and_max_rank(lambda: _b200(dtype))
    _compress_chain(lambda: _b200(dtype))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", DTYPES)
def test_gpu_reference_mps_contraction(dtype):
    _mps_matches_full_wavefunction(lambda: _b200(dtype), TOL[np.dtype(dtype)])
    _inorder_with_bond_merging(lambda: _b200(dtype), TOL[np.dtype(dtype)])


def _mps_dsl_stream():
    """The MPS contraction of prep(3,1)+GHZ-3 recorded as a .tl stream (decompose commands)."""
    from picoquant_jl_b200.host.backends import TensorStore
    circ = create_simple_preparation_circuit(3, 1).compose(create_ghz_preparation_circuit(3))
    dsl = DSLBackend()
    tn = convert_circuit_to_network(circ, dsl, decompose=True)
    add_input(tn, "000")
    mps_nodes = contract_mps_tensor_network_circuit(tn)
    calculate_mps_amplitudes(tn, mps_nodes)
    assert any(l.startswith("decompose ") for l in dsl.text().splitlines())
    ob = OracleBackend(np.complex128)
    tn = convert_circuit_to_network(circ, ob, decompose=True)
    add_input(tn, "000")
    full_wavefunction_contraction(tn, "vector")
    return dsl, np.array(ob.load_tensor_data("result")), TensorStore


def test_dsl_interpreter_with_decompose_oracle():
    from oracle.interactive import execute_dsl
    dsl, ref, TensorStore = _mps_dsl_stream()
    out = TensorStore()
    execute_dsl(dsl.text(), dsl.store, np.complex128, output_store=out)
    assert rel_l2(out.read("result"), ref) < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", DTYPES)
def test_gpu_execute_dsl_with_decompose(dtype):
    """execute_dsl_file on the device for a stream with `decompose` commands (interpreted,
    chi is only known at run time) and for one without (compiled program)."""
    dsl, ref, TensorStore = _mps_dsl_stream()
    b = _b200(dtype)
    out = TensorStore()
    b.execute_dsl(dsl.text(), dsl.store, out)
    assert rel_l2(out.read("result"), ref) < 10 * TOL[np.dtype(dtype)]
    plain = DSLBackend()
    tn = convert_circuit_to_network(create_qft_circuit(4), plain)
    add_input(tn, "0101")
    full_wavefunction_contraction(tn, "vector")
    ob = OracleBackend(np.complex128)
    tn = convert_circuit_to_network(create_qft_circuit(4), ob)
    add_input(tn, "0101")
    full_wavefunction_contraction(tn, "vector")
    out = TensorStore()
    b.execute_dsl(plain.text(), plain.store, out)
    assert rel_l2(out.read("result"), ob.load_tensor_data("result")) < 10 * TOL[np.dtype(dtype)]


@pytest.mark.gpu
def test_gpu_decompose_large_bond():
    """A 512 x 384 bond matrix (2 chi x 2 chi of an MPS with chi ~ 200): the multi-launch
    Jacobi path at a size where the sweep count matters; singular values to 1e-10."""
    rng = np.random.default_rng(8)
    t = rand_tensor(rng, (2, 256, 192, 2), np.complex128)
    b = _b200(np.complex128)
    b.save_tensor_data("T", t)
    chi = b.decompose_tensor("T", [1, 2], [3, 4], max_rank=100, left_label="L", right_label="R")
    assert chi == 100
    check_split(t, [1, 2], [3, 4], b.load_tensor_data("L"), b.load_tensor_data("R"), chi, 1e-13, 100,
                1e-10)
