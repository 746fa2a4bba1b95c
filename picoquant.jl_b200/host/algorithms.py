"""Circuit generators used as bench / test inputs.  Restates the gate sequences
of ``src/algorithms.jl``: ``create_qft_circuit`` (:14-32),
``create_simple_preparation_circuit`` (:41-72),
``create_ghz_preparation_circuit`` (:79-89) and the Google-style random quantum
circuit ``create_RQC`` (:118-277).

The reference draws random numbers from Julia's ``MersenneTwister``; that
stream cannot be reproduced here, so random choices come from
``numpy.random.default_rng(seed)`` and the seed is always part of the workload
name.  Gate order, patterns and the "first single-qubit gate is T, never repeat"
rule are identical.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np

from .circuit import Circuit


def create_qft_circuit(n: int, m: int = -1) -> Circuit:
    """``src/algorithms.jl:14-32``: for i=1..n: h(i-1); cu1(pi/2^(j-i+1), j, i-1)
    for j=i..n-1 (only when j-i < m); barrier; then swap(i-1, n-i) for
    i=1..floor(n/2)."""
    circ = Circuit(n)
    m = n if m == -1 else m
    for i in range(1, n + 1):
        circ.h(i - 1)
        for j in range(i, n):
            if j - i < m:
                circ.cu1(math.pi / (2 ** (j - i + 1)), j, i - 1)
        circ.barrier()
    for i in range(1, n // 2 + 1):
        circ.swap(i - 1, n - i)
    return circ


def create_simple_preparation_circuit(qubits: int, depth: int,
                                      seed: Optional[int] = None) -> Circuit:
    """``src/algorithms.jl:41-72``: H layer, then ``depth`` layers of random
    u3 on every qubit followed by a brick of CX gates (start qubit alternates)."""
    rng = np.random.default_rng(seed)
    circ = Circuit(qubits)
    for q in range(1, qubits + 1):
        circ.h(q - 1)
    circ.barrier()
    for d in range(1, depth + 1):
        params = rng.random((qubits, 3))
        for q in range(1, qubits + 1):
            circ.u3(params[q - 1, 0], params[q - 1, 1], params[q - 1, 2], q - 1)
        start_qubit = ((d + 1) % 2) + 1
        for q in range(start_qubit, qubits, 2):
            circ.cx(q - 1, q)
        circ.barrier()
    return circ


def create_ghz_preparation_circuit(qubits: int) -> Circuit:
    """``src/algorithms.jl:79-89``."""
    circ = Circuit(qubits)
    circ.h(0)
    for q in range(1, qubits):
        circ.cx(q - 1, q)
    return circ


def rqc_patterns(rows: int, cols: int) -> Dict[int, List[List[List[int]]]]:
    """``patterns`` (``src/algorithms.jl:148-166``) with the bounds filter of
    ``out_of_bounds`` (:173-177).  Keys 1..8; every entry is a list of qubit
    pairs ``[[i, j], [u, v]]`` in 1-based grid coordinates."""
    gp: Dict[int, List[List[List[int]]]] = {}
    key = 1
    N = max(rows, cols)
    for v in (1, -1):
        for shift in (0, 1):
            for flip in (False, True):
                pairs = []
                # Julia comprehension: `for j in 1:2:N for i in ...` -> j outer, i inner
                for j in range(1, N + 1, 2):
                    for i in range(1 + ((j + shift) // 2) % 2, N + 1, 2):
                        pairs.append([[i, j], [i, j + v]])
                if flip:
                    pairs = [[p[0][::-1], p[1][::-1]] for p in pairs]
                kept = []
                for (i, j), (u, w) in pairs:
                    if 0 < i <= rows and 0 < u <= rows and 0 < j <= cols and 0 < w <= cols:
                        kept.append([[i, j], [u, w]])
                gp[key] = kept
                key += 1
    return gp


_RQC_PATTERN_ORDER = [3, 1, 6, 8, 5, 7, 2, 4]


def create_RQC(rows: int, cols: int, depth: int, seed: Optional[int] = None, *,
               use_iswap: bool = False, final_Hadamard_layer: bool = False) -> Circuit:
    """``create_RQC`` (``src/algorithms.jl:199-277``).  Qubit (i, j) (1-based
    grid coordinates) is circuit qubit ``i + (j-1)*rows - 1``; single-qubit gate
    ids 1:T, 2:rx(pi/2), 3:ry(pi/2) (:118-121, :220-224)."""
    rng = np.random.default_rng(seed)
    circ = Circuit(rows * cols)
    next_gate = -np.ones((rows, cols), dtype=np.int64)

    def q(i: int, j: int) -> int:
        return i + (j - 1) * rows - 1

    def single(gate_id: int, i: int, j: int) -> None:
        if gate_id == 1:
            circ.t(q(i, j))
        elif gate_id == 2:
            circ.rx(math.pi / 2, q(i, j))
        else:
            circ.ry(math.pi / 2, q(i, j))

    for i in range(1, rows + 1):
        for j in range(1, cols + 1):
            circ.h(q(i, j))

    gate_patterns = rqc_patterns(rows, cols)
    for d in range(depth):
        key = _RQC_PATTERN_ORDER[d % 8]
        pat = gate_patterns[key]
        if len(pat) != 0:
            for (i, j), (u, w) in pat:
                if use_iswap:
                    circ.iswap(q(i, j), q(u, w))
                else:
                    circ.cz(q(i, j), q(u, w))
            qubits_hit = {tuple(p) for pair in pat for p in pair}
            for i in range(1, rows + 1):
                for j in range(1, cols + 1):
                    if (i, j) not in qubits_hit:
                        gate = int(next_gate[i - 1, j - 1])
                        if gate > 0:
                            # random_gate! (:133-142)
                            nxt = (gate + int(rng.integers(0, 2))) % 3 + 1
                            next_gate[i - 1, j - 1] = -nxt
                            single(gate, i, j)
            for (i, j), (u, w) in pat:
                next_gate[i - 1, j - 1] = abs(next_gate[i - 1, j - 1])
                next_gate[u - 1, w - 1] = abs(next_gate[u - 1, w - 1])
        circ.barrier()

    if final_Hadamard_layer:
        for i in range(1, rows + 1):
            for j in range(1, cols + 1):
                circ.h(q(i, j))
    return circ
