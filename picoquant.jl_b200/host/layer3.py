"""Tensor-network circuit data structure (host side).

Restates ``src/layer3.jl`` of the reference: ``Node`` (:23-30), ``Edge``
(:43-53), ``TensorNetworkCircuit`` (:56-118), backend forwarders (:124-134),
``new_label!`` (:145-148), ``add_gate!`` (:158-249), ``add_input!`` (:257-284),
``add_output!`` (:292-313), neighbour/edge accessors (:324-455), circuit ingest
(:543-573), JSON (:584-707), ``decompose_gate!`` (:715-739) and the built-in
gate tensors (:746-802).

Everything here is integer/label bookkeeping plus tiny host-side gate algebra;
all tensor *data* goes to ``network.backend`` through the nine backend calls.
Qubit numbers passed to ``add_gate`` are 1-based like the reference.
"""
from __future__ import annotations

import json
from typing import Dict, List, Optional, Sequence

import numpy as np

from .backends import AbstractBackend
from .circuit import Circuit, gate_matrix

def _asf(a):
    """Fortran-contiguous view/copy that keeps 0-d arrays 0-d."""
    return np.asarray(a, order="F")



class Node:
    """``src/layer3.jl:23-30``: index labels, dims and the data label."""

    __slots__ = ("indices", "dims", "data_label")

    def __init__(self, indices: Optional[Sequence[str]] = None,
                 dims: Optional[Sequence[int]] = None, data_label: str = "") -> None:
        self.indices: List[str] = list(indices) if indices is not None else []
        self.dims: List[int] = [int(d) for d in dims] if dims is not None else []
        self.data_label = data_label

    def __repr__(self) -> str:
        return "Node(%r, %r, %r)" % (self.indices, self.dims, self.data_label)


class Edge:
    """``src/layer3.jl:43-53``."""

    __slots__ = ("src", "dst", "qubit", "virtual")

    def __init__(self, src: Optional[str] = None, dst: Optional[str] = None,
                 qubit: Optional[int] = None, virtual: bool = False) -> None:
        self.src = src
        self.dst = dst
        self.qubit = qubit
        self.virtual = bool(virtual)

    def __repr__(self) -> str:
        return "Edge(%r, %r, %r, %r)" % (self.src, self.dst, self.qubit, self.virtual)


class TensorNetworkCircuit:
    """``src/layer3.jl:56-118``.  ``nodes`` and ``edges`` are insertion-ordered
    (Python dicts == the reference's ``OrderedDict``); ``node_layers`` is an
    unordered ``Dict`` in the reference, so every consumer here sorts it
    explicitly (see ``layer2._layer_nodes``)."""

    def __init__(self, qubits: int, backend: AbstractBackend) -> None:
        if backend is None:
            raise ValueError("a backend is required")
        self.backend = backend
        self.number_qubits = int(qubits)
        index_labels = ["index_%d" % i for i in range(1, qubits + 1)]
        self.edges: Dict[str, Edge] = {}
        for i, lab in enumerate(index_labels, start=1):
            self.edges[lab] = Edge(None, None, i, False)
        self.input_qubits: List[str] = list(index_labels)
        self.output_qubits: List[str] = list(index_labels)
        self.nodes: Dict[str, Node] = {}
        self.qubit_ordering: List[int] = list(range(1, qubits + 1))
        self.counters: Dict[str, int] = {"index": qubits, "node": 0, "layer": 0}
        self.node_layers: Dict[str, int] = {}

    # -- backend forwarders (src/layer3.jl:124-134) -------------------------
    def save_tensor_data(self, *a, **k): return self.backend.save_tensor_data(*a, **k)
    def load_tensor_data(self, *a, **k): return self.backend.load_tensor_data(*a, **k)
    def contract_tensors(self, *a, **k): return self.backend.contract_tensors(*a, **k)
    def save_output(self, *a, **k): return self.backend.save_output(*a, **k)
    def reshape_tensor(self, *a, **k): return self.backend.reshape_tensor(*a, **k)
    def permute_tensor(self, *a, **k): return self.backend.permute_tensor(*a, **k)
    def decompose_tensor(self, *a, **k): return self.backend.decompose_tensor(*a, **k)
    def delete_tensor(self, *a, **k): return self.backend.delete_tensor(*a, **k)
    def view_tensor(self, *a, **k): return self.backend.view_tensor(*a, **k)


def new_label(network: TensorNetworkCircuit, label_str: str) -> str:
    """``src/layer3.jl:145-148``."""
    network.counters[label_str] += 1
    return "%s_%d" % (label_str, network.counters[label_str])


def _fortran(a) -> np.ndarray:
    return _asf(np.array(a))


def decompose_gate(gate_data, threshold: float = 1e-15):
    """``src/layer3.jl:715-739``: split a (2,2,2,2) gate tensor with indices
    [in1,in2,out1,out2] into B[in1,out1,v] and C[v,in2,out2] by an SVD of the
    (1,3|2,4) matricisation, keeping singular values > ``threshold`` (absolute)
    and putting sqrt(S) on both sides."""
    g = _fortran(gate_data)
    if g.ndim != 4:
        raise ValueError("decompose_gate needs a rank-4 tensor")
    left_positions, right_positions = [0, 2], [1, 3]
    dims = g.shape
    left_dims = [dims[x] for x in left_positions]
    right_dims = [dims[x] for x in right_positions]
    a = np.transpose(g, left_positions + right_positions)
    a = np.reshape(a, (int(np.prod(left_dims)), int(np.prod(right_dims))), order="F")
    u, s, vt = np.linalg.svd(a, full_matrices=False)
    chi = int(np.sum(s > threshold))
    sq = np.sqrt(s[:chi])
    b = np.reshape(u[:, :chi] * sq[None, :], tuple(left_dims) + (chi,), order="F")
    c = np.reshape(sq[:, None] * vt[:chi, :], (chi,) + tuple(right_dims), order="F")
    return _fortran(b), _fortran(c)


def _remap_wires(network: TensorNetworkCircuit, node_label: str,
                 input_index: str, output_index: str, qubit: int) -> None:
    # shared tail of both branches of add_gate! (layer3.jl:200-212 / 229-245)
    network.edges[output_index] = Edge(node_label, network.edges[input_index].dst, qubit)
    dst = network.edges[input_index].dst
    if dst is not None:
        out_node = network.nodes[dst]
        for i in range(len(out_node.indices)):
            if out_node.indices[i] == input_index:
                out_node.indices[i] = output_index
    network.edges[input_index].dst = node_label


def add_gate(network: TensorNetworkCircuit, gate_data, target_qubits: Sequence[int],
             decompose: bool = False):
    """``src/layer3.jl:158-249``.  ``gate_data`` has axes
    [in_1..in_k, out_1..out_k]; ``target_qubits`` are 1-based.  Returns the new
    node label, or the two labels when a 2-qubit gate is SVD-split."""
    gate_data = _fortran(gate_data)
    target_qubits = [int(q) for q in target_qubits]
    n = len(target_qubits)
    input_indices = [network.output_qubits[q - 1] for q in target_qubits]
    output_indices = [new_label(network, "index") for _ in range(n)]
    for i, q in enumerate(target_qubits):
        network.output_qubits[q - 1] = output_indices[i]

    network.counters["layer"] += 1
    layer = network.counters["layer"]
    if decompose and n == 2:
        gates_data = decompose_gate(gate_data)
        virtual_index = new_label(network, "index")
        node_labels = [new_label(network, "node") for _ in range(2)]
        network.edges[virtual_index] = Edge(node_labels[0], node_labels[1], None, True)
        for i in range(2):
            node_label = node_labels[i]
            if i == 0:
                indices = [input_indices[i], output_indices[i], virtual_index]
            else:
                indices = [virtual_index, input_indices[i], output_indices[i]]
            network.nodes[node_label] = Node(indices, list(gates_data[i].shape), node_label)
            network.node_layers[node_label] = layer
            network.save_tensor_data(node_label, gates_data[i])
            _remap_wires(network, node_label, input_indices[i], output_indices[i],
                         target_qubits[i])
        return node_labels

    node_label = new_label(network, "node")
    network.nodes[node_label] = Node(input_indices + output_indices,
                                     list(gate_data.shape), node_label)
    network.node_layers[node_label] = layer
    network.save_tensor_data(node_label, gate_data)
    for k in range(n):
        _remap_wires(network, node_label, input_indices[k], output_indices[k],
                     target_qubits[k])
    return node_label


_INPUT_CAPS = {
    "0": np.array([1.0, 0.0]),
    "1": np.array([0.0, 1.0]),
    "+": np.array([1.0, 1.0]) / np.sqrt(2.0),
    "-": np.array([1.0, -1.0]) / np.sqrt(2.0),
}


def add_input(network: TensorNetworkCircuit, config: str) -> None:
    """``src/layer3.jl:257-284`` (layer 0 caps; existing caps are left alone)."""
    assert len(config) == network.number_qubits
    for input_index, ch in zip(network.input_qubits, config):
        if network.edges[input_index].src is None:
            node_label = new_label(network, "node")
            node_data = _INPUT_CAPS[ch]
            network.nodes[node_label] = Node([input_index], [2], node_label)
            network.node_layers[node_label] = 0
            network.save_tensor_data(node_label, node_data)
            network.edges[input_index].src = node_label
        # else: the reference only logs "Input node already exists"


def add_output(network: TensorNetworkCircuit, config: str) -> None:
    """``src/layer3.jl:292-313`` (layer -1 caps, placed per qubit_ordering)."""
    assert len(config) == network.number_qubits
    for i in range(network.number_qubits):
        qubit_pos = network.qubit_ordering[i]
        output_index, ch = network.output_qubits[qubit_pos - 1], config[i]
        if network.edges[output_index].dst is None:
            node_label = new_label(network, "node")
            node_data = np.array([1.0, 0.0]) if ch == "0" else np.array([0.0, 1.0])
            network.nodes[node_label] = Node([output_index], [2], node_label)
            network.node_layers[node_label] = -1
            network.save_tensor_data(node_label, node_data)
            network.edges[output_index].dst = node_label


# -- accessors (src/layer3.jl:324-455) --------------------------------------

def edges(network: TensorNetworkCircuit) -> Dict[str, Edge]:
    return network.edges


def inneighbours(network, node_label):
    out = []
    for index in network.nodes[node_label].indices:
        e = network.edges[index]
        if e.src is not None and e.src != node_label and not e.virtual:
            out.append(e.src)
    return out


def outneighbours(network, node_label):
    out = []
    for index in network.nodes[node_label].indices:
        e = network.edges[index]
        if e.dst is not None and e.dst != node_label and not e.virtual:
            out.append(e.dst)
    return out


def virtualneighbours(network, node_label):
    out = []
    for index in network.nodes[node_label].indices:
        e = network.edges[index]
        if e.dst is not None and e.dst != node_label and e.virtual:
            out.append(e.dst)
        elif e.src is not None and e.src != node_label and e.virtual:
            out.append(e.src)
    return out


def neighbours(network, node_label):
    return (inneighbours(network, node_label) + outneighbours(network, node_label)
            + virtualneighbours(network, node_label))


def inedges(network, node_label):
    return [x for x in network.nodes[node_label].indices
            if not network.edges[x].virtual and network.edges[x].dst == node_label]


def outedges(network, node_label):
    return [x for x in network.nodes[node_label].indices
            if not network.edges[x].virtual and network.edges[x].src == node_label]


def virtualedges(network, node_label):
    return [x for x in network.nodes[node_label].indices if network.edges[x].virtual]


def getedge(network, edge_label):
    return network.edges[edge_label]


def getnode(network, node_label):
    return network.nodes[node_label]


# -- circuit ingest (src/layer3.jl:543-573) -----------------------------------

def gate_data_from_matrix(u: np.ndarray) -> np.ndarray:
    """``src/layer3.jl:565-569``: ``reshape(permutedims(U,(2,1)), 2,...,2)`` in
    column-major order, i.e. T[in_1..in_k, out_1..out_k] = U[out, in] with the
    first target qubit the least-significant bit."""
    u = np.asarray(u)
    k = int(round(np.log2(u.shape[0])))
    return np.reshape(_asf(u.T), (2,) * (2 * k), order="F")


def convert_circuit_to_network(circ: Circuit, backend: AbstractBackend, *,
                               decompose: bool = False, transpile: bool = False,
                               couplings=None) -> TensorNetworkCircuit:
    """``convert_qiskit_circ_to_network`` (``src/layer3.jl:543-573``) for the
    qiskit-free ``Circuit``.  ``transpile=True`` routes the circuit onto the coupling map
    with the restated BasicSwap pass (``circuit.transpile_circuit``) and records where each
    logical qubit ended up in ``qubit_ordering`` (``layer3.jl:553-560``)."""
    qubit_ordering = None
    if transpile:
        from .circuit import transpile_circuit
        circ, qubit_ordering = transpile_circuit(circ, couplings)
    tng = TensorNetworkCircuit(circ.n_qubits, backend)
    if qubit_ordering is not None:
        tng.qubit_ordering[:] = qubit_ordering
    for name, params, qubits in circ.data:
        if name == "barrier":
            continue
        data = gate_data_from_matrix(gate_matrix(name, params))
        add_gate(tng, data, [q + 1 for q in qubits], decompose=decompose)
    return tng


convert_qiskit_circ_to_network = convert_circuit_to_network


# -- built-in gate tensors (src/layer3.jl:746-802) ---------------------------

def _mat(rows):
    return np.array(rows, dtype=np.complex128)


GATE_TENSORS = {
    "I": _mat([[1, 0], [0, 1]]),
    "X": _mat([[0, 1], [1, 0]]),
    # stored untransposed in the reference (layer3.jl:750)
    "Y": _mat([[0, -1j], [1j, 0]]),
    "Z": _mat([[1, 0], [0, -1]]),
    "H": _mat([[1, 1], [1, -1]]) / np.sqrt(2.0),
    "S": _mat([[1, 0], [0, 1j]]),
    "T": _mat([[1, 0], [0, (1 + 1j) / np.sqrt(2.0)]]),
    "CX": np.reshape(_asf(_mat([[1, 0, 0, 0], [0, 0, 0, 1],
                                             [0, 0, 1, 0], [0, 1, 0, 0]])),
                     (2, 2, 2, 2), order="F"),
    "CZ": np.reshape(_asf(_mat([[1, 0, 0, 0], [0, 1, 0, 0],
                                             [0, 0, 1, 0], [0, 0, 0, -1]])),
                     (2, 2, 2, 2), order="F"),
    "SWAP": np.reshape(_asf(_mat([[1, 0, 0, 0], [0, 0, 1, 0],
                                               [0, 1, 0, 0], [0, 0, 0, 1]])),
                       (2, 2, 2, 2), order="F"),
}


def gate_tensor(gate: str) -> np.ndarray:
    """``src/layer3.jl:787-802``."""
    if gate not in GATE_TENSORS:
        raise ValueError("Invalid input to gate_tensor: %s" % gate)
    return _asf(GATE_TENSORS[gate].copy())


# -- JSON (src/layer3.jl:584-707) ---------------------------------------------

def to_dict(network: TensorNetworkCircuit) -> dict:
    top = {"number_qubits": network.number_qubits}
    top["edges"] = {k: {"src": e.src, "dst": e.dst, "virtual": e.virtual, "qubit": e.qubit}
                    for k, e in network.edges.items()}
    top["nodes"] = {k: {"indices": list(n.indices), "dims": [str(d) for d in n.dims],
                        "data_label": n.data_label}
                    for k, n in network.nodes.items()}
    top["input_qubits"] = list(network.input_qubits)
    top["output_qubits"] = list(network.output_qubits)
    top["qubit_ordering"] = list(network.qubit_ordering)
    top["node_layers"] = {k: network.node_layers[k]
                          for k in sorted(network.node_layers, key=_label_number)}
    return top


def _label_number(label: str) -> int:
    return int(label.rsplit("_", 1)[-1])


def network_from_dict(d: dict, backend: AbstractBackend) -> TensorNetworkCircuit:
    net = TensorNetworkCircuit(d["number_qubits"], backend)
    net.counters = {"index": 0, "node": 0}
    net.edges = {}
    for k, v in d["edges"].items():
        net.counters["index"] = max(net.counters["index"], _label_number(k))
        net.edges[k] = Edge(v["src"], v["dst"], v.get("qubit"), v.get("virtual", False))
    net.nodes = {}
    for k, v in d["nodes"].items():
        net.counters["node"] = max(net.counters["node"], _label_number(k))
        dims = [int(x) for x in v["dims"]] if "dims" in v else [int(x) for x in v.get("data_dims", [])]
        net.nodes[k] = Node(v["indices"], dims, v.get("data_label", k))
    net.input_qubits = list(d["input_qubits"])
    net.output_qubits = list(d["output_qubits"])
    net.qubit_ordering = list(d.get("qubit_ordering", range(1, net.number_qubits + 1)))
    net.node_layers = dict(d.get("node_layers", {}))
    return net


def to_json(network: TensorNetworkCircuit, indent: int = 0) -> str:
    d = to_dict(network)
    return json.dumps(d, separators=(",", ":")) if indent == 0 else json.dumps(d, indent=indent)


def network_from_json(json_str: str, backend: AbstractBackend) -> TensorNetworkCircuit:
    return network_from_dict(json.loads(json_str), backend)
