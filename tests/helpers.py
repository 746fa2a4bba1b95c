"""Shared helpers for the test-suite (the analogue of test/test_utils.jl)."""
import json
import os

import numpy as np

import picoquant_jl_b200  # noqa: F401  (registers the package)
from picoquant_jl_b200.host import (add_input, add_output, contract_network,
                                    convert_circuit_to_network,
                                    full_wavefunction_contraction)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_l2(x, ref):
    x = np.asarray(x).ravel(order="F")
    ref = np.asarray(ref).ravel(order="F")
    den = np.linalg.norm(ref)
    return float(np.linalg.norm(x - ref) / (den if den > 0 else 1.0))


TOL = {np.dtype(np.complex128): 1e-10, np.dtype(np.complex64): 1e-5}


def switch_endianness(vec):
    """test/test_utils.jl:13-18"""
    n = int(round(np.log2(len(vec))))
    t = np.reshape(vec, (2,) * n, order="F")
    return np.reshape(np.transpose(t, list(range(n - 1, -1, -1))), 2 ** n, order="F")


def statevector(circ, backend, input_config=None, **kw):
    """get_statevector_using_picoquant (test/test_utils.jl:43-55), little-endian."""
    tn = convert_circuit_to_network(circ, backend, **kw)
    add_input(tn, input_config or "0" * circ.n_qubits)
    full_wavefunction_contraction(tn, "vector")
    return np.array(backend.load_tensor_data("result"))


def amplitude(circ, backend, plan_fn, input_config=None, output_config=None, **kw):
    tn = convert_circuit_to_network(circ, backend, **kw)
    n = circ.n_qubits
    add_input(tn, input_config or "0" * n)
    add_output(tn, output_config or "0" * n)
    contract_network(tn, plan_fn(tn))
    return np.array(backend.load_tensor_data("result"))


def load_golden_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def golden_qasm(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return f.read()
