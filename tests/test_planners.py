"""The reference's own planner API, restated in host/planner.py (plans are inputs to the hot
path): netcon_contraction! / bgreedy_contraction! leave one node (test/layer2_tests.jl:145-203)
and give the right amplitude; the netcon restatement (subset DP standing in for
TensorOperations.optimaltree) is optimal against brute force; the compression driver
contract_tensor_network_circuit_with_compression! reproduces the state.  CPU only."""
import itertools
import random

import numpy as np

from helpers import rel_l2
from oracle.interactive import OracleBackend
from picoquant_jl_b200.host import (TensorNetworkCircuit, add_input, add_output, bgreedy,
                                    bgreedy_contraction, calculate_mps_amplitudes,
                                    contract_tensor_network_circuit_with_compression,
                                    convert_circuit_to_network, create_ghz_preparation_circuit,
                                    create_qft_circuit, full_wavefunction_contraction,
                                    load_qasm_as_circuit, netcon, netcon_contraction, plan_cost)

QASM = """OPENQASM 2.0;
          include "qelib1.inc";
          qreg q[3];
          h q[0];
          cx q[0],q[1];
          cx q[0],q[2];
          """


def _network():
    b = OracleBackend(np.complex128)
    tn = convert_circuit_to_network(load_qasm_as_circuit(QASM), b)
    add_input(tn, "000")
    add_output(tn, "000")
    return tn, b


def test_reference_netcon_and_bgreedy_contraction():
    for fn in (netcon_contraction, lambda tn: bgreedy_contraction(tn, rng=random.Random(3))):
        tn, b = _network()
        fn(tn)
        assert len(tn.nodes) == 1
        assert abs(b.load_tensor_data("result") - 1 / np.sqrt(2)) < 1e-12


def test_bgreedy_plan_shape_and_reproducibility():
    tn, _ = _network()
    p1, t1, s1 = bgreedy(tn, 1.0, 1.0, rng=random.Random(7))
    p2, t2, s2 = bgreedy(tn, 1.0, 1.0, rng=random.Random(7))
    assert p1 == p2 and t1 == t2 and s1 == s2
    assert len(p1) == len(tn.nodes) - 1 and plan_cost(tn, p1)["remaining"] == 1
    best, cost = bgreedy(tn, 1.0, 1.0, 20, rng=random.Random(7))
    assert cost <= t1 and plan_cost(tn, best)["remaining"] == 1
    # labels of intermediates follow convert_tree_to_plan / contract_pair!: node_{counter+k}
    made = {"node_%d" % (tn.counters["node"] + k + 1) for k in range(len(p1))}
    used = {x for pair in p1 for x in pair}
    assert used - set(tn.nodes) <= made


def _brute_force_best(tn):
    """Minimum total M*N*K over all binary contraction trees (for a handful of tensors):
    exhaustive recursion over which pair to contract next, memoised on the set of groups."""
    start = {frozenset([k]): (list(v.indices), list(v.dims)) for k, v in tn.nodes.items()}
    memo = {}

    def rec(groups):
        if len(groups) == 1:
            return 0
        key = frozenset(groups)
        if key in memo:
            return memo[key]
        best = None
        for a, b in itertools.combinations(list(groups), 2):
            ia, da = groups[a]
            ib, db = groups[b]
            dim = dict(zip(ia + ib, da + db))
            rest = [x for x in ia if x not in ib] + [x for x in ib if x not in ia]
            cost = 1
            for x in set(ia + ib):
                cost *= dim[x]
            g = {k: v for k, v in groups.items() if k not in (a, b)}
            g[a | b] = (rest, [dim[x] for x in rest])
            c = cost + rec(g)
            if best is None or c < best:
                best = c
        memo[key] = best
        return best

    return rec(start)


def test_netcon_is_optimal_against_brute_force():
    rng = random.Random(11)
    for trial in range(6):
        b = OracleBackend(np.complex128)
        tn = TensorNetworkCircuit(2, b)
        circ = create_qft_circuit(2)
        tn = convert_circuit_to_network(circ, b)
        bits = "".join(rng.choice("01") for _ in range(2))
        add_input(tn, bits)
        if trial % 2:
            add_output(tn, "10")
        assert len(tn.nodes) <= 8
        plan = netcon(tn)
        assert plan_cost(tn, plan)["remaining"] == 1
        assert plan_cost(tn, plan)["macs"] == _brute_force_best(tn)


def test_netcon_refuses_large_networks():
    b = OracleBackend(np.complex128)
    tn = convert_circuit_to_network(create_qft_circuit(5), b)
    add_input(tn, "00000")
    try:
        netcon(tn)
    except ValueError:
        return
    raise AssertionError("expected ValueError for %d tensors" % len(tn.nodes))


def test_compression_driver_reproduces_the_state():
    """src/layer2.jl:657-702 on GHZ-4 (nearest-neighbour gates): the saved site tensors
    contract to (|0000> + |1111>)/sqrt(2)."""
    n = 4
    b = OracleBackend(np.complex128)
    tn = convert_circuit_to_network(create_ghz_preparation_circuit(n), b, decompose=True)
    add_input(tn, "0" * n)
    sites = contract_tensor_network_circuit_with_compression(tn)
    assert len(sites) == n and all(b.load_tensor_data(s) is not None for s in sites)
    calculate_mps_amplitudes(tn, sites)
    ref = np.zeros(2 ** n, dtype=np.complex128)
    ref[[0, -1]] = 1 / np.sqrt(2)
    assert rel_l2(b.load_tensor_data("result"), ref) < 1e-10
