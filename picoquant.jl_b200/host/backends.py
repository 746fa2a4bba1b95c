"""Backend interface of PicoQuant, restated for the Python host mirror.

Reference: ``src/backends.jl:3-57`` (``AbstractBackend``, ``Metrics``,
``record_compute_costs!``, ``record_memory_costs!``) and the nine generic
functions exported at ``src/backends.jl:62-66`` and forwarded from
``TensorNetworkCircuit`` at ``src/layer3.jl:124-134``.

Conventions kept from the reference: labels are plain strings (Julia Symbols),
every index / axis / range value crossing this interface is **1-based**, tensor
data is **column-major**.  Julia's ``f!`` names lose the bang in Python.
"""
from __future__ import annotations

import json
from typing import Dict, List, Optional, Sequence

import numpy as np

def _asf(a):
    """Fortran-contiguous view/copy that keeps 0-d arrays 0-d."""
    return np.asarray(a, order="F")



class Metrics:
    """``src/backends.jl:8-20`` -- counters updated by ``contract_pair``."""

    def __init__(self) -> None:
        self.max_tensor_size = 0
        self.total_space_allocated = 0
        self.flops = 0

    def as_tuple(self):
        return (self.max_tensor_size, self.total_space_allocated, self.flops)

    def __repr__(self) -> str:
        return ("Metrics(max_tensor_size=%d, total_space_allocated=%d, flops=%d)"
                % self.as_tuple())


def _prod(values: Sequence[int]) -> int:
    out = 1
    for v in values:
        out *= int(v)
    return out


class AbstractBackend:
    """``src/backends.jl:3``.  Concrete backends carry ``metrics`` and implement
    the nine functions below (same names and argument meaning as the
    reference)."""

    metrics: Metrics

    def save_tensor_data(self, tensor_label: str, tensor_data) -> None:
        raise NotImplementedError

    def load_tensor_data(self, tensor_label: str):
        raise NotImplementedError

    def contract_tensors(self, A_label: str, A_ncon_indices: Sequence[int],
                         B_label: str, B_ncon_indices: Sequence[int],
                         C_label: str) -> None:
        raise NotImplementedError

    def save_output(self, node: str, name: str = "result") -> None:
        raise NotImplementedError

    def reshape_tensor(self, tensor: str, groups: Sequence[Sequence[int]]) -> None:
        raise NotImplementedError

    def permute_tensor(self, tensor: str, axes: Sequence[int]) -> None:
        raise NotImplementedError

    def decompose_tensor(self, tensor: str, left_positions: Sequence[int],
                         right_positions: Sequence[int], *, threshold: float = 1e-13,
                         max_rank: int = 0, left_label: str, right_label: str) -> int:
        raise NotImplementedError

    def delete_tensor(self, tensor_label: str) -> None:
        raise NotImplementedError

    def view_tensor(self, view_node: str, node: str, bond_idx: int,
                    bond_range: Sequence[int]) -> None:
        raise NotImplementedError


def record_memory_costs(backend: AbstractBackend, tensor_memory: int) -> None:
    """``src/backends.jl:50-57``."""
    backend.metrics.total_space_allocated += tensor_memory
    if tensor_memory > backend.metrics.max_tensor_size:
        backend.metrics.max_tensor_size = tensor_memory


def record_compute_costs(backend: AbstractBackend, open_dims: Sequence[int],
                         contracted_dims: Sequence[int]):
    """``src/backends.jl:30-41``: space = prod(open dims) (empty product = 1),
    flops += space * prod(contracted dims)."""
    space_cost = _prod(open_dims)
    record_memory_costs(backend, space_cost)
    flops_cost = space_cost * _prod(contracted_dims)
    backend.metrics.flops += flops_cost
    return flops_cost, space_cost


class TensorStore:
    """Stand-in for the HDF5 tensor file used by the reference DSL backend
    (``src/backends/dsl.jl:65-79``, ``src/layer1.jl:22-64``).  libhdf5 is not
    available in this image, so datasets are kept in a dict keyed by the same
    dataset names and can be written to / read from an ``.npz`` side-car."""

    def __init__(self) -> None:
        self.data: Dict[str, np.ndarray] = {}

    def write(self, key: str, array) -> None:
        self.data[key] = _asf(np.array(array))

    def read(self, key: str) -> Optional[np.ndarray]:
        return self.data.get(key)

    def exists(self, key: str) -> bool:
        return key in self.data

    def save_npz(self, path: str) -> None:
        np.savez(path, **self.data)

    @classmethod
    def load_npz(cls, path: str) -> "TensorStore":
        store = cls()
        with np.load(path) as f:
            for k in f.files:
                store.data[k] = _asf(f[k])
        return store


def _join(values: Sequence[int]) -> str:
    # Julia's ``join(v, ",")``; a UnitRange joins element-wise (dsl.jl:212).
    return ",".join(str(int(v)) for v in values)


def _julia_float(x: float) -> str:
    """String interpolation of a Float64 the way Julia prints it (shortest
    round-trip repr, ``1.0e-13`` style exponents)."""
    x = float(x)
    r = repr(x)
    if "e" in r:
        mant, exp = r.split("e")
        if "." not in mant:
            mant += ".0"
        return "%se%d" % (mant, int(exp))
    return r


class DSLBackend(AbstractBackend):
    """Deferred backend that records the textual command stream
    (``src/backends/dsl.jl:14-214``).  The ``.tl`` stream is the bit-exact
    record of contraction-plan and slice indexing; it is what the plan-parity
    tests compare and what ``pq_program_compile`` consumes on the device side.
    Commands are kept in ``self.commands`` and, when a filename is given, also
    appended to that file like the reference does."""

    def __init__(self, dsl: Optional[str] = None, tensor_data: str = "tensor_data.h5",
                 output: str = "", overwrite: bool = False,
                 store: Optional[TensorStore] = None) -> None:
        self.dsl_filename = dsl
        self.tensor_data_filename = tensor_data
        self.output_data_filename = output if output != "" else tensor_data
        self.metrics = Metrics()
        self.commands: List[str] = []
        self.store = store if store is not None else TensorStore()
        self.output_store = self.store if output == "" else TensorStore()
        if dsl is not None:
            import os
            if overwrite or not os.path.isfile(dsl):
                open(dsl, "w").close()

    # dsl.jl:49-53
    def push(self, instruction: str) -> None:
        self.commands.append(instruction)
        if self.dsl_filename is not None:
            with open(self.dsl_filename, "a") as io:
                io.write(instruction + "\n")

    def text(self) -> str:
        return "".join(c + "\n" for c in self.commands)

    # dsl.jl:65-79
    def save_tensor_data(self, tensor_label, tensor_data):
        self.store.write(str(tensor_label), tensor_data)
        self.push("tensor %s %s" % (tensor_label, tensor_label))

    # dsl.jl:87-102
    def load_tensor_data(self, tensor_label):
        store = self.output_store if tensor_label == "result" else self.store
        return store.read(str(tensor_label))

    # dsl.jl:113-127
    def contract_tensors(self, A_label, A_ncon_indices, B_label, B_ncon_indices, C_label):
        self.push("ncon %s %s %s %s %s" % (C_label, A_label, _join(A_ncon_indices),
                                           B_label, _join(B_ncon_indices)))
        self.push("del %s" % A_label)
        self.push("del %s" % B_label)

    # dsl.jl:135-138
    def save_output(self, node, name="result"):
        self.push("save %s %s %s" % (node, self.output_data_filename, name))

    # dsl.jl:147-152
    def reshape_tensor(self, tensor, groups):
        self.push("reshape %s %s" % (tensor, ";".join(_join(g) for g in groups)))

    # dsl.jl:160-165
    def permute_tensor(self, tensor, axes):
        self.push("permute %s %s" % (tensor, _join(axes)))

    # dsl.jl:181-195 (returns 0: rank unknown until run time)
    def decompose_tensor(self, tensor, left_positions, right_positions, *,
                         threshold=1e-13, max_rank=0, left_label, right_label):
        self.push('decompose %s %s %s %s %s {"threshold":%s, "max_rank":%d}'
                  % (tensor, left_label, _join(left_positions), right_label,
                     _join(right_positions), _julia_float(threshold), max_rank))
        return 0

    # dsl.jl:202-204
    def delete_tensor(self, tensor_label):
        self.push("del %s" % tensor_label)

    # dsl.jl:211-214
    def view_tensor(self, view_node, node, bond_idx, bond_range):
        self.push("view %s %s %d %s" % (view_node, node, bond_idx, _join(bond_range)))


def parse_dsl(text: str):
    """Tokenise a ``.tl`` stream the way ``execute_dsl_file`` does
    (``src/layer1.jl:222-312``): whitespace-split, first token is the command.
    Returns a list of ``(command, args_dict)`` tuples with integer lists parsed.
    """
    ops = []
    for line in text.splitlines():
        tok = line.split()
        if not tok:
            continue
        cmd = tok[0]
        if cmd == "ncon":
            C, A, a_idx, B, b_idx = tok[1:6]
            ops.append(("ncon", dict(C=C, A=A, B=B,
                                     a_idx=[int(x) for x in a_idx.split(",")],
                                     b_idx=[int(x) for x in b_idx.split(",")])))
        elif cmd == "del":
            ops.append(("del", dict(t=tok[1])))
        elif cmd == "tensor":
            ops.append(("tensor", dict(t=tok[1], key=tok[2])))
        elif cmd == "save":
            ops.append(("save", dict(t=tok[1], file=tok[2], key=tok[3])))
        elif cmd == "reshape":
            groups = [[int(y) for y in x.split(",")] for x in tok[2].split(";")]
            ops.append(("reshape", dict(t=tok[1], groups=groups)))
        elif cmd == "permute":
            ops.append(("permute", dict(t=tok[1], axes=[int(x) for x in tok[2].split(",")])))
        elif cmd == "decompose":
            A, Bl, b_idx, Cl, c_idx = tok[1:6]
            options = json.loads("".join(tok[6:]))
            ops.append(("decompose", dict(t=A, left=Bl, right=Cl,
                                          left_idx=[int(x) for x in b_idx.split(",")],
                                          right_idx=[int(x) for x in c_idx.split(",")],
                                          options=options)))
        elif cmd == "view":
            ops.append(("view", dict(v=tok[1], t=tok[2], axis=int(tok[3]),
                                     idx=[int(x) for x in tok[4].split(",")])))
        # unknown commands are silently ignored, as in the reference interpreter
    return ops
