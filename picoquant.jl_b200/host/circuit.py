"""Minimal circuit container, gate matrices and an OPENQASM 2.0 subset parser.

The reference ingests circuits through qiskit 0.19.2 via PyCall
(``src/layer3.jl:467-573``); qiskit is not available in this image, so the
pieces of it that the hot path depends on are restated here:

* the gate unitaries of ``qelib1.inc`` in qiskit's little-endian convention
  (first listed qubit = least-significant bit of the matrix index), which is
  what ``qi.Operator(gate).data`` returns at ``src/layer3.jl:567``;
* a ``Circuit`` object with a ``.data`` list of ``(name, params, qubits)``
  (0-based qubits, like qiskit), barriers included so that they can be skipped
  the way ``src/layer3.jl:563`` does;
* ``load_qasm_as_circuit`` / ``load_qasm_as_circuit_from_file``
  (``src/layer3.jl:467-484``).

Gate *tensors* are derived from these matrices in ``layer3.gate_data_from_matrix``
following ``src/layer3.jl:565-569`` (``reshape(transpose(U), 2, ..., 2)``).
"""
from __future__ import annotations

import ast
import cmath
import math
import operator
import os
import re
from typing import List, Sequence, Tuple

import numpy as np

SQRT2 = math.sqrt(2.0)


def _u3(theta: float, phi: float, lam: float) -> np.ndarray:
    c, s = math.cos(theta / 2), math.sin(theta / 2)
    return np.array([[c, -cmath.exp(1j * lam) * s],
                     [cmath.exp(1j * phi) * s, cmath.exp(1j * (phi + lam)) * c]],
                    dtype=np.complex128)


def _controlled(u: np.ndarray) -> np.ndarray:
    """Controlled-U with the control as the FIRST listed qubit (= LSB in
    qiskit's little-endian matrix): rows/cols with bit0 == 1 carry U."""
    m = np.eye(4, dtype=np.complex128)
    m[1, 1], m[1, 3], m[3, 1], m[3, 3] = u[0, 0], u[0, 1], u[1, 0], u[1, 1]
    return m


_X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
_Y = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
_Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)
_H = np.array([[1, 1], [1, -1]], dtype=np.complex128) / SQRT2


def gate_matrix(name: str, params: Sequence[float] = ()) -> np.ndarray:
    """Unitary of a named gate (qiskit 0.19.2 definitions, little-endian)."""
    p = [float(x) for x in params]
    n = name.lower()
    if n in ("id", "i", "iden"):
        return np.eye(2, dtype=np.complex128)
    if n == "x":
        return _X.copy()
    if n == "y":
        return _Y.copy()
    if n == "z":
        return _Z.copy()
    if n == "h":
        return _H.copy()
    if n == "s":
        return np.diag([1, 1j]).astype(np.complex128)
    if n == "sdg":
        return np.diag([1, -1j]).astype(np.complex128)
    if n == "t":
        return np.diag([1, cmath.exp(1j * math.pi / 4)]).astype(np.complex128)
    if n == "tdg":
        return np.diag([1, cmath.exp(-1j * math.pi / 4)]).astype(np.complex128)
    if n == "u1":
        return np.diag([1, cmath.exp(1j * p[0])]).astype(np.complex128)
    if n == "u2":
        return _u3(math.pi / 2, p[0], p[1])
    if n in ("u3", "u"):
        return _u3(p[0], p[1], p[2])
    if n == "rx":
        c, s = math.cos(p[0] / 2), math.sin(p[0] / 2)
        return np.array([[c, -1j * s], [-1j * s, c]], dtype=np.complex128)
    if n == "ry":
        c, s = math.cos(p[0] / 2), math.sin(p[0] / 2)
        return np.array([[c, -s], [s, c]], dtype=np.complex128)
    if n == "rz":
        return np.diag([cmath.exp(-0.5j * p[0]), cmath.exp(0.5j * p[0])]).astype(np.complex128)
    if n in ("cx", "cnot"):
        return _controlled(_X)
    if n == "cy":
        return _controlled(_Y)
    if n == "cz":
        return _controlled(_Z)
    if n == "ch":
        return _controlled(_H)
    if n == "cu1":
        return _controlled(np.diag([1, cmath.exp(1j * p[0])]))
    if n == "crz":
        return _controlled(gate_matrix("rz", p))
    if n == "cu3":
        return _controlled(_u3(p[0], p[1], p[2]))
    if n == "swap":
        return np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]],
                        dtype=np.complex128)
    if n == "iswap":
        return np.array([[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]],
                        dtype=np.complex128)
    if n in ("ccx", "toffoli"):
        # controls = first two listed qubits (bits 0 and 1), target = bit 2
        m = np.eye(8, dtype=np.complex128)
        m[3, 3] = m[7, 7] = 0
        m[3, 7] = m[7, 3] = 1
        return m
    raise ValueError("unsupported gate: %r" % name)


GATE_ARITY = {"ccx": 3, "toffoli": 3}
for _g in ("cx", "cnot", "cy", "cz", "ch", "cu1", "crz", "cu3", "swap", "iswap"):
    GATE_ARITY[_g] = 2


class Circuit:
    """Ordered list of gate applications on ``n_qubits`` qubits (0-based, like
    qiskit's ``QuantumCircuit``).  ``data`` entries are
    ``(name, params_tuple, qubits_tuple)``; ``("barrier", (), qubits)`` entries
    are kept and skipped at ingest, as the reference does."""

    def __init__(self, n_qubits: int) -> None:
        self.n_qubits = int(n_qubits)
        self.data: List[Tuple[str, Tuple[float, ...], Tuple[int, ...]]] = []

    # -- generic append ----------------------------------------------------
    def append(self, name: str, params: Sequence[float], qubits: Sequence[int]) -> "Circuit":
        qubits = tuple(int(q) for q in qubits)
        for q in qubits:
            if not 0 <= q < self.n_qubits:
                raise IndexError("qubit %d out of range" % q)
        if len(set(qubits)) != len(qubits):
            raise ValueError("duplicate qubits in gate %s%r" % (name, qubits))
        self.data.append((name.lower(), tuple(float(x) for x in params), qubits))
        return self

    def barrier(self, *qubits: int) -> "Circuit":
        qs = qubits if qubits else tuple(range(self.n_qubits))
        self.data.append(("barrier", (), tuple(qs)))
        return self

    # -- qiskit-style helpers used by the circuit generators ----------------
    def h(self, q): return self.append("h", (), (q,))
    def x(self, q): return self.append("x", (), (q,))
    def y(self, q): return self.append("y", (), (q,))
    def z(self, q): return self.append("z", (), (q,))
    def s(self, q): return self.append("s", (), (q,))
    def t(self, q): return self.append("t", (), (q,))
    def rx(self, theta, q): return self.append("rx", (theta,), (q,))
    def ry(self, theta, q): return self.append("ry", (theta,), (q,))
    def rz(self, phi, q): return self.append("rz", (phi,), (q,))
    def u1(self, lam, q): return self.append("u1", (lam,), (q,))
    def u2(self, phi, lam, q): return self.append("u2", (phi, lam), (q,))
    def u3(self, theta, phi, lam, q): return self.append("u3", (theta, phi, lam), (q,))
    def cx(self, c, t): return self.append("cx", (), (c, t))
    def cz(self, c, t): return self.append("cz", (), (c, t))
    def cu1(self, lam, c, t): return self.append("cu1", (lam,), (c, t))
    def swap(self, a, b): return self.append("swap", (), (a, b))
    def iswap(self, a, b): return self.append("iswap", (), (a, b))

    def compose(self, other: "Circuit") -> "Circuit":
        """qiskit's ``compose``/``combine`` for same-width circuits."""
        if other.n_qubits != self.n_qubits:
            raise ValueError("circuit width mismatch")
        out = Circuit(self.n_qubits)
        out.data = list(self.data) + list(other.data)
        return out

    combine = compose

    def gates(self):
        """Non-barrier operations, in order."""
        return [g for g in self.data if g[0] != "barrier"]

    def qasm(self) -> str:
        lines = ["OPENQASM 2.0;", 'include "qelib1.inc";', "qreg q[%d];" % self.n_qubits]
        for name, params, qubits in self.data:
            args = ",".join("q[%d]" % q for q in qubits)
            if params:
                lines.append("%s(%s) %s;" % (name, ",".join(repr(p) for p in params), args))
            else:
                lines.append("%s %s;" % (name, args))
        return "\n".join(lines) + "\n"

    def simulate(self, input_config: str = None) -> np.ndarray:
        """Dense little-endian state-vector simulation (known-answer tests only;
        stands in for qiskit Aer in test/algorithms_tests.jl).  Input caps follow
        ``add_input!`` ('0', '1', '+', '-')."""
        n = self.n_qubits
        caps = {"0": [1, 0], "1": [0, 1], "+": [2 ** -0.5, 2 ** -0.5], "-": [2 ** -0.5, -2 ** -0.5]}
        cfg = input_config or "0" * n
        psi = np.ones(1, dtype=np.complex128)
        for ch in cfg:            # qubit 0 is the fastest (least significant) axis
            psi = np.kron(np.array(caps[ch], dtype=np.complex128), psi)
        psi = psi.reshape((2,) * n)   # C-order: axis 0 <-> qubit n-1, axis n-1 <-> qubit 0
        for name, params, qubits in self.gates():
            k = len(qubits)
            g = gate_matrix(name, params).reshape((2,) * (2 * k))  # [out_{k-1}..out_0, in_{k-1}..in_0]
            axes = [n - 1 - q for q in reversed(qubits)]           # state axes of q_{k-1}..q_0
            psi = np.tensordot(g, psi, axes=(list(range(k, 2 * k)), axes))
            psi = np.moveaxis(psi, list(range(k)), axes)
        return psi.reshape(-1)

    def to_matrix(self) -> np.ndarray:
        """Dense unitary (little-endian), for tiny known-answer tests only."""
        n = self.n_qubits
        dim = 1 << n
        u = np.eye(dim, dtype=np.complex128)
        for name, params, qubits in self.gates():
            g = gate_matrix(name, params)
            k = len(qubits)
            full = np.zeros((dim, dim), dtype=np.complex128)
            for col in range(dim):
                sub_in = sum(((col >> q) & 1) << i for i, q in enumerate(qubits))
                rest = col
                for q in qubits:
                    rest &= ~(1 << q)
                for sub_out in range(1 << k):
                    amp = g[sub_out, sub_in]
                    if amp != 0:
                        row = rest
                        for i, q in enumerate(qubits):
                            row |= ((sub_out >> i) & 1) << q
                        full[row, col] += amp
            u = full @ u
        return u


# ---------------------------------------------------------------------------
# OPENQASM 2.0 subset
# ---------------------------------------------------------------------------

_BINOPS = {ast.Add: operator.add, ast.Sub: operator.sub, ast.Mult: operator.mul,
           ast.Div: operator.truediv, ast.Pow: operator.pow}
_FUNCS = {"sin": math.sin, "cos": math.cos, "tan": math.tan, "exp": math.exp,
          "ln": math.log, "sqrt": math.sqrt}


def _eval_param(expr: str) -> float:
    node = ast.parse(expr.replace("^", "**"), mode="eval").body

    def ev(n):
        if isinstance(n, ast.Constant) and isinstance(n.value, (int, float)):
            return float(n.value)
        if isinstance(n, ast.Name) and n.id == "pi":
            return math.pi
        if isinstance(n, ast.BinOp) and type(n.op) in _BINOPS:
            return _BINOPS[type(n.op)](ev(n.left), ev(n.right))
        if isinstance(n, ast.UnaryOp) and isinstance(n.op, (ast.USub, ast.UAdd)):
            v = ev(n.operand)
            return -v if isinstance(n.op, ast.USub) else v
        if isinstance(n, ast.Call) and isinstance(n.func, ast.Name) and n.func.id in _FUNCS:
            return _FUNCS[n.func.id](*[ev(a) for a in n.args])
        raise ValueError("unsupported expression in qasm parameter: %r" % expr)

    return ev(node)


def _split_top_level(s: str) -> List[str]:
    out, depth, cur = [], 0, []
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            out.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
    if cur:
        out.append("".join(cur))
    return [x.strip() for x in out if x.strip()]


_STMT = re.compile(r"^\s*([A-Za-z_][A-Za-z0-9_]*)\s*(?:\((.*)\))?\s*(.*)$", re.S)
_ARG = re.compile(r"^([A-Za-z_][A-Za-z0-9_]*)\s*(?:\[\s*(\d+)\s*\])?$")


def load_qasm_as_circuit(qasm_str: str) -> Circuit:
    """``src/layer3.jl:481-484`` restated for the qelib1 subset used by the
    reference's fixtures (``examples/*.qasm``).  Multiple ``qreg``s are laid out
    consecutively; ``creg``/``measure`` are ignored; a gate applied to a whole
    register is broadcast."""
    text = re.sub(r"//[^\n]*", "", qasm_str)
    stmts = [s.strip() for s in text.split(";") if s.strip()]
    qregs = {}
    total = 0
    ops = []
    for s in stmts:
        if s.startswith("OPENQASM") or s.startswith("include"):
            continue
        m = _STMT.match(s)
        if not m:
            raise ValueError("cannot parse qasm statement: %r" % s)
        name, params, args = m.group(1), m.group(2), m.group(3)
        if name == "qreg":
            am = _ARG.match(args.strip())
            qregs[am.group(1)] = (total, int(am.group(2)))
            total += int(am.group(2))
            continue
        if name in ("creg", "measure", "reset"):
            continue
        if name in ("gate", "opaque", "if"):
            raise ValueError("unsupported qasm construct: %r" % name)
        ops.append((name, params, args))
    circ = Circuit(total)
    for name, params, args in ops:
        pvals = tuple(_eval_param(p) for p in _split_top_level(params)) if params else ()
        targets = []
        for a in _split_top_level(args):
            am = _ARG.match(a)
            if not am or am.group(1) not in qregs:
                raise ValueError("bad qasm argument %r" % a)
            base, size = qregs[am.group(1)]
            if am.group(2) is None:
                targets.append([base + i for i in range(size)])
            else:
                targets.append([base + int(am.group(2))])
        if name == "barrier":
            circ.barrier(*[q for t in targets for q in t])
            continue
        width = max(len(t) for t in targets)
        for i in range(width):
            qs = [t[i] if len(t) > 1 else t[0] for t in targets]
            circ.append(name, pvals, qs)
    return circ


def load_qasm_as_circuit_from_file(qasm_path: str):
    """``src/layer3.jl:467-474`` (returns ``False`` when the file is missing)."""
    if os.path.isfile(qasm_path):
        with open(qasm_path) as f:
            return load_qasm_as_circuit(f.read())
    return False


def transpile_circuit(circ: Circuit, couplings=None):
    """``transpile_circuit`` (``src/layer3.jl:493-527``): make every two-qubit gate act on
    coupled qubits.  The reference runs qiskit's ``BasicSwap`` pass; its published algorithm
    is restated here: walk the gates in order under a logical -> physical layout; when a
    two-qubit gate's qubits are further apart than one coupling, take the shortest
    (undirected) path between them and swap the FIRST qubit along it until the two are
    neighbours -- ``for k in range(len(path) - 2): swap(path[k], path[k+1])`` -- updating the
    layout (nothing is swapped back).  ``couplings`` defaults to the line
    ``[[i-1, i] for i = 1:n]`` like the reference (pairs naming a qubit >= n are ignored).
    Returns ``(circuit on physical qubits, qubit_ordering)`` with
    ``qubit_ordering[c] = 1 + physical qubit holding logical qubit c`` -- what the reference
    reads back from the measurements it appended (``layer3.jl:516-524``); its own test pins
    ``[2, 1, 3]`` for ``h q0; cx q0,q2`` (``test/layer3_tests.jl:89-102``)."""
    n = circ.n_qubits
    if couplings is None:
        couplings = [[i - 1, i] for i in range(1, n + 1)]
    adj = {q: set() for q in range(n)}
    for a, b in couplings:
        if 0 <= a < n and 0 <= b < n and a != b:
            adj[a].add(b)
            adj[b].add(a)

    def shortest_path(src: int, dst: int):
        prev = {src: None}
        frontier = [src]
        while frontier and dst not in prev:
            nxt = []
            for u in frontier:
                for v in sorted(adj[u]):
                    if v not in prev:
                        prev[v] = u
                        nxt.append(v)
            frontier = nxt
        if dst not in prev:
            raise ValueError("qubits %d and %d are not connected by the coupling map" % (src, dst))
        path = [dst]
        while prev[path[-1]] is not None:
            path.append(prev[path[-1]])
        return path[::-1]

    layout = list(range(n))          # logical -> physical
    out = Circuit(n)
    for name, params, qubits in circ.data:
        if name == "barrier":
            out.data.append((name, params, tuple(layout[q] for q in qubits)))
            continue
        if len(qubits) > 2:
            raise ValueError("transpile_circuit handles one- and two-qubit gates only")
        if len(qubits) == 2:
            path = shortest_path(layout[qubits[0]], layout[qubits[1]])
            for k in range(len(path) - 2):
                a, b = path[k], path[k + 1]
                out.swap(a, b)
                la, lb = layout.index(a), layout.index(b)
                layout[la], layout[lb] = b, a
        out.append(name, params, [layout[q] for q in qubits])
    return out, [layout[c] + 1 for c in range(n)]
