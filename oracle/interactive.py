"""In-memory CPU backend -- the reference's ``InteractiveBackend{Array{T}}``
restated on NumPy (TEST INFRASTRUCTURE; see ``oracle/__init__.py``).

Follows ``src/backends/interactive.jl`` one-for-one: ``save_tensor_data``
(:32-36, converts to the backend element type), ``load_tensor_data`` (:44-49,
``None`` when absent), ``contract_tensors`` (:60-75, stores C then deletes A and
B), ``save_output`` (:84-88, alias, no copy), ``reshape_tensor`` (:97-102),
``permute_tensor`` (:111-115), ``delete_tensor!`` (:159-161, missing label is
not an error), ``view_tensor!`` (:169-172).
"""
from __future__ import annotations

import os
import sys
from typing import Dict

import numpy as np

_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _root not in sys.path:
    sys.path.insert(0, _root)

from picoquant_jl_b200.host.backends import AbstractBackend, Metrics  # noqa: E402

from . import layer1  # noqa: E402

def _asf(a):
    """Fortran-contiguous view/copy that keeps 0-d arrays 0-d."""
    return np.asarray(a, order="F")



class OracleBackend(AbstractBackend):
    """``InteractiveBackend{Array{dtype}}``; ``dtype`` defaults to complex64
    like the reference's default constructor (interactive.jl:19-22)."""

    def __init__(self, dtype=np.complex64) -> None:
        self.dtype = np.dtype(dtype)
        self.tensors: Dict[str, np.ndarray] = {}
        self.metrics = Metrics()

    def save_tensor_data(self, tensor_label, tensor_data):
        self.tensors[tensor_label] = _asf(
            np.array(tensor_data).astype(self.dtype, copy=False))

    def load_tensor_data(self, tensor_label):
        return self.tensors.get(tensor_label)

    def contract_tensors(self, A_label, A_ncon_indices, B_label, B_ncon_indices, C_label):
        A = self.tensors[A_label]
        B = self.tensors[B_label]
        C = layer1.contract_tensors((A, B), (list(A_ncon_indices), list(B_ncon_indices)))
        self.save_tensor_data(C_label, C)
        self.delete_tensor(A_label)
        self.delete_tensor(B_label)

    def save_output(self, node, name="result"):
        self.tensors[name] = self.tensors[node]

    def reshape_tensor(self, tensor, groups):
        dims = self.tensors[tensor].shape
        new_dims = []
        for g in groups:
            d = 1
            for x in g:
                d *= dims[x - 1]
            new_dims.append(d)
        self.tensors[tensor] = layer1.reshape_tensor(self.tensors[tensor], new_dims)

    def permute_tensor(self, tensor, axes):
        self.tensors[tensor] = layer1.permute_tensor(self.tensors[tensor], list(axes))

    def decompose_tensor(self, tensor, left_positions, right_positions, *, threshold=1e-13,
                         max_rank=0, left_label, right_label):
        """``src/backends/interactive.jl:130-152``."""
        B, C, chi = layer1.decompose_tensor(self.tensors[tensor], list(left_positions),
                                            list(right_positions), threshold, max_rank)
        self.tensors[left_label] = B
        self.tensors[right_label] = C
        if tensor not in (left_label, right_label):
            self.delete_tensor(tensor)
        return chi

    def delete_tensor(self, tensor_label):
        self.tensors.pop(tensor_label, None)

    def view_tensor(self, view_node, node, bond_idx, bond_range):
        self.tensors[view_node] = layer1.tensor_view(self.tensors[node], bond_idx,
                                                     list(bond_range))


def execute_dsl(text: str, store, dtype=np.complex64, output_store=None) -> Dict[str, np.ndarray]:
    """Interpreter half of ``execute_dsl_file`` (``src/layer1.jl:211-315``) over
    an in-memory ``.tl`` text and a ``TensorStore`` (HDF5 stand-in).  ``save``
    writes into ``output_store`` (default: the same store) under the given
    dataset name.  Returns the final tensor dictionary."""
    from picoquant_jl_b200.host.backends import parse_dsl

    dtype = np.dtype(dtype)
    out_store = output_store if output_store is not None else store
    tensors: Dict[str, np.ndarray] = {}
    for cmd, a in parse_dsl(text):
        if cmd == "ncon":
            tensors[a["C"]] = layer1.contract_tensors(
                (tensors[a["A"]], tensors[a["B"]]), (a["a_idx"], a["b_idx"])).astype(dtype, copy=False)
        elif cmd == "del":
            tensors.pop(a["t"], None)
        elif cmd == "tensor":
            tensors[a["t"]] = _asf(store.read(a["key"]).astype(dtype))
        elif cmd == "save":
            out_store.write(a["key"], tensors[a["t"]])
        elif cmd == "reshape":
            dims = tensors[a["t"]].shape
            new_dims = [int(np.prod([dims[y - 1] for y in g])) for g in a["groups"]]
            tensors[a["t"]] = layer1.reshape_tensor(tensors[a["t"]], new_dims)
        elif cmd == "permute":
            tensors[a["t"]] = layer1.permute_tensor(tensors[a["t"]], a["axes"])
        elif cmd == "view":
            tensors[a["v"]] = layer1.tensor_view(tensors[a["t"]], a["axis"], a["idx"])
        elif cmd == "decompose":   # src/layer1.jl:286-300 (the source tensor is kept)
            B, C, _ = layer1.decompose_tensor(tensors[a["t"]], a["left_idx"], a["right_idx"],
                                              float(a["options"].get("threshold", 1e-13)),
                                              int(a["options"].get("max_rank", 0)))
            tensors[a["left"]] = B
            tensors[a["right"]] = C
    return tensors
