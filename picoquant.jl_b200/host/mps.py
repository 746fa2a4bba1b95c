"""``MPSState`` (``src/mps.jl``): array-style access to single amplitudes of a matrix
product state left behind by ``contract_mps_tensor_network_circuit!``.

Like the reference's, this is a *host-side* utility: the constructor downloads the site
tensors once (``load_tensor_data``), splits each along its physical (output-qubit) axis, and
``state["0101…"]`` / ``state[i1, i2, …]`` contracts the selected matrices left to right with
the same index bookkeeping the drivers use (``sort_indices`` / ``create_ncon_indices``).  The
per-amplitude work is a chain of tiny matrix products on host arrays, exactly what
``src/mps.jl:86-100`` does with ``ncon`` on CPU arrays; nothing here touches the device hot
path.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

from .layer2 import create_ncon_indices, sort_indices
from .layer3 import Node, TensorNetworkCircuit


def _ncon_pair(a: np.ndarray, a_idx: Sequence[int], b: np.ndarray, b_idx: Sequence[int]) -> np.ndarray:
    """``ncon([a, b], (a_idx, b_idx))`` for one pair: positive labels are contracted,
    negative labels are open and ordered -1, -2, …"""
    con = [l for l in a_idx if l > 0]
    out = np.tensordot(a, b, axes=([list(a_idx).index(l) for l in con],
                                   [list(b_idx).index(l) for l in con]))
    open_labels = [l for l in a_idx if l < 0] + [l for l in b_idx if l < 0]
    order = sorted(range(len(open_labels)), key=lambda k: -open_labels[k])
    return np.transpose(out, order) if order else out


class MPSState:
    """``src/mps.jl:9-56``.  ``dtype`` defaults to ComplexF32 like the reference."""

    def __init__(self, network: TensorNetworkCircuit, mps_nodes: Sequence[str],
                 dtype=np.complex64) -> None:
        n = len(mps_nodes)
        self.nodes: List[Node] = []
        self.data_tensors: List[List[np.ndarray]] = []
        output_positions = []
        outputs = set(network.output_qubits)
        for i, label in enumerate(mps_nodes, start=1):
            node = network.nodes[label]
            pos = next(k for k, x in enumerate(node.indices) if x in outputs)
            output_positions.append(pos + 1)
            others = [k for k in range(len(node.indices)) if k != pos]
            tensor = np.transpose(np.asarray(network.load_tensor_data(label)), [pos] + others)
            self.nodes.append(Node([node.indices[k] for k in others],
                                   [node.dims[k] for k in others], "n_%d" % i))
            self.data_tensors.append([np.asarray(tensor[x], dtype=dtype) for x in range(2)])
        # src/mps.jl:50-53 builds `ordering[e] = i` by zipping `output_positions` -- the AXIS
        # position of the physical index inside each site tensor -- with `qubit_ordering`,
        # where the SITE number is evidently meant (its own test, GHZ-5 with amplitudes
        # 00000 / 11111 / 10101, cannot tell the two apart).  The mirror uses the site number,
        # so that state[bits] equals the corresponding entry of calculate_mps_amplitudes!.
        self.ordering = [0] * n
        for site, e in enumerate(network.qubit_ordering, start=1):
            self.ordering[e - 1] = site
        self.output_positions = output_positions
        self.n = n

    @property
    def shape(self):           # size(a)
        return (2,) * self.n

    def __len__(self) -> int:  # length(a)
        return 2 ** self.n

    def amplitude(self, index: Sequence[int]):
        """``getindex(a, i...)`` with 1-based entries (1 -> |0>, 2 -> |1>)."""
        if len(index) != self.n:
            raise IndexError("MPSState expects %d indices" % self.n)
        conf = lambda x: 0 if x % 2 == 1 else 1
        n1 = self.nodes[0]
        data = self.data_tensors[0][conf(index[self.ordering[0] - 1])]
        for idx in range(1, self.n):
            n2 = self.nodes[idx]
            common, remaining = sort_indices(n1, n2)
            a_idx, b_idx = create_ncon_indices(n1, n2, common, remaining)
            n1 = Node(remaining, [1] * len(remaining), "n1")
            data = _ncon_pair(data, a_idx, self.data_tensors[idx][conf(index[self.ordering[idx] - 1])],
                              b_idx)
        return complex(np.asarray(data).reshape(-1)[0])

    def __getitem__(self, key):
        if isinstance(key, str):     # src/mps.jl:107-109
            key = [1 if c == "0" else 2 for c in key]
        elif isinstance(key, int):
            key = [key]
        return self.amplitude(list(key))
