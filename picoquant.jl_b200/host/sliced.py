"""Sliced contraction driver: the flow of ``examples/dist_slicing_example.jl`` /
``test/layer2_tests.jl:419-455`` organised for a device backend.

The reference rebuilds the whole network for every partition (uploading every
gate again), slices it, contracts it and finally reduces the partial results.
All partitions have identical shapes and an identical plan -- only the index
picked by each ``view`` differs -- so here the host mirror is run **once**
against a recording ``DSLBackend``; the resulting ``.tl`` stream is compiled by
``pq_program_compile`` and replayed per partition with the ``view`` start
indices of that partition (``multi_index_partition``, bit-exact with
``src/layer2/slicing.jl:12-26``).  Partials are accumulated on the device and
summed across GPUs by one NCCL all-reduce.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from .backends import DSLBackend, parse_dsl
from .circuit import Circuit
from .layer2 import contract_network, full_wavefunction_contraction
from .layer3 import TensorNetworkCircuit, add_input, add_output, convert_circuit_to_network
from .slicing import (multi_index_partition, partition_network_on_virtual_bonds,
                      slice_tensor_network)

PlanFn = Callable[[TensorNetworkCircuit, Sequence[str]], List[List[str]]]


class SlicedRecording:
    """The command stream of one partition plus what is needed to re-target it."""

    def __init__(self, text: str, store, bond_labels: List[str], bond_dims: List[int],
                 number_partitions: int, partition: int, metrics) -> None:
        self.text = text
        self.store = store
        self.bond_labels = bond_labels
        self.bond_dims = bond_dims
        self.number_partitions = number_partitions
        self.partition = partition
        self.metrics = metrics

    def view_starts(self, partition: int) -> List[int]:
        """1-based start index of every ``view`` command for ``partition``: both
        end nodes of sliced bond j are restricted to bond value j
        (``slice_tensor_network``, slicing.jl:100-110)."""
        values = multi_index_partition(tuple(self.bond_dims), self.number_partitions, partition)
        out: List[int] = []
        for v in values:
            out.extend([v, v])
        return out

    def text_for(self, partition: int) -> str:
        """The stream the host mirror would emit for another partition (used by
        the tests to prove that only the ``view`` indices differ)."""
        starts = self.view_starts(partition)
        lines, k = [], 0
        for line in self.text.splitlines():
            tok = line.split()
            if tok and tok[0] == "view":
                n = len(tok[4].split(","))
                tok[4] = ",".join(str(starts[k] + i) for i in range(n))
                k += 1
                line = " ".join(tok)
            lines.append(line)
        return "".join(l + "\n" for l in lines)


def record_sliced_contraction(circ: Circuit, number_partitions: int, partition: int = 1, *,
                              plan_fn: Optional[PlanFn] = None, input_config: Optional[str] = None,
                              output_config: Optional[str] = None,
                              output_shape="") -> SlicedRecording:
    """Runs the reference flow for one partition against a recording backend:
    ``convert_qiskit_circ_to_network(decompose=true)`` -> ``add_input!`` ->
    (``add_output!``) -> ``partition_network_on_virtual_bonds`` ->
    ``slice_tensor_network`` -> ``contract_network!(plan)`` or, without a
    ``plan_fn``, ``full_wavefunction_contraction!``.  ``number_partitions=0``
    records the unsliced contraction."""
    dsl = DSLBackend()
    tn = convert_circuit_to_network(circ, dsl, decompose=True)
    n = circ.n_qubits
    add_input(tn, input_config or "0" * n)
    if output_config is not None:
        add_output(tn, output_config)
    bond_labels: List[str] = []
    bond_dims: List[int] = []
    if number_partitions > 0:
        labels, values = partition_network_on_virtual_bonds(tn, number_partitions, partition)
        bond_labels = list(labels)
        for lab in bond_labels:
            e = tn.edges[lab]
            data = dsl.load_tensor_data(e.src)
            pos = tn.nodes[e.src].indices.index(lab)
            bond_dims.append(int(data.shape[pos]))
        slice_tensor_network(tn, labels, values)
    if plan_fn is not None:
        plan = plan_fn(tn, bond_labels)
        contract_network(tn, plan, output_shape)
    else:
        full_wavefunction_contraction(tn, output_shape)
    return SlicedRecording(dsl.text(), dsl.store, bond_labels, bond_dims, number_partitions,
                           partition, dsl.metrics)


def partitions_of_rank(number_partitions: int, rank: int, world: int) -> List[int]:
    """Contiguous block of 1-based partition ids owned by ``rank`` (the
    reference runs exactly one partition per MPI rank,
    dist_slicing_example.jl:7-8; here a GPU owns P/world of them)."""
    per = number_partitions // world
    extra = number_partitions % world
    start = rank * per + min(rank, extra)
    count = per + (1 if rank < extra else 0)
    return list(range(start + 1, start + count + 1))


class SlicedContraction:
    """Compiled, replayable sliced contraction on one ``B200Backend``."""

    def __init__(self, backend, recording: SlicedRecording, upload: bool = True) -> None:
        self.backend = backend
        self.rec = recording
        self._upload_args = None
        if upload:
            self.upload()
        self.program = backend.compile_program(recording.text)
        assert self.program.num_views == 2 * len(recording.bond_labels)

    def upload(self) -> int:
        """Host -> device copy of every leaf tensor named by a ``tensor``
        command (the HDF5 file of the reference DSL flow; the per-rank re-upload of
        examples/dist_slicing_example.jl:22-27) through ONE ``pq_save_tensors`` call:
        the host arrays are packed into a pinned staging block, copied with a single
        H2D transfer and scattered by one launch.  Returns the bytes copied.  The
        ctypes argument arrays are prepared once; the packing and the copy happen on
        every call."""
        if self._upload_args is None:
            items = [(a["key"], self.rec.store.read(a["key"]))
                     for cmd, a in parse_dsl(self.rec.text) if cmd == "tensor"]
            self._upload_args = self.backend.prepare_save_batch(items)
        args, nbytes = self._upload_args
        self.backend.save_prepared_batch(args)
        return nbytes

    def run(self, partitions: Sequence[int], accumulate_into: str = "partial_sum",
            hoist: bool = False, lanes: int = 1) -> None:
        """Contracts the given partitions and accumulates their results on the device.

        ``hoist=False`` (default) replays the whole command stream for every partition,
        like the reference flow does.  ``hoist=True`` executes the slice-invariant part of
        the stream once per call (``pq_program_prepare``) and only the slice-dependent part
        per partition -- same result, less work.  ``lanes > 1`` keeps that many partitions
        in flight on the device (``pq_program_run_slices``); the partial sums are still added
        in partition order, so the result is bit-identical."""
        self.program.set_hoist(hoist)
        if hoist:
            self.program.prepare()
        if lanes > 1 and self.rec.bond_labels and len(partitions) > 1:
            self.program.run_slices([self.rec.view_starts(p) for p in partitions],
                                    accumulate_into, lanes)
            return
        for p in partitions:
            self.program.run(self.rec.view_starts(p) if self.rec.bond_labels else None,
                             accumulate_into)

    def result(self, label: str = "partial_sum") -> np.ndarray:
        return self.backend.load_tensor_data(label)
