"""Slicing a tensor network on its virtual bonds.  Restates
``src/layer2/slicing.jl``: ``multi_index_partition`` (:12-26),
``partition_network_on_virtual_bonds`` (:37-57), ``replace_with_view!``
(:67-93) and ``slice_tensor_network`` (:100-110).

Pure integer work: partition ids, bond values and ranges are 1-based and must
match the reference bit for bit.
"""
from __future__ import annotations

import logging
from typing import List, Sequence, Tuple

from .layer3 import Node, TensorNetworkCircuit, new_label

_log = logging.getLogger(__name__)


def multi_index_partition(dims: Sequence[int], number_partitions: int,
                          partition: int) -> Tuple[int, ...]:
    """``src/layer2/slicing.jl:12-26``.  Finds the shortest prefix of ``dims``
    whose product equals ``number_partitions`` and returns
    ``CartesianIndices(dims[1:k])[partition]`` -- the column-major unravel of
    the 1-based ``partition`` (first bond fastest), as a tuple of 1-based bond
    values.  Quirks kept: the loop never inspects the last dim, and a product
    that overshoots only logs an error and keeps going (App. D.1), so e.g.
    ``number_partitions=1`` returns an index over *all* bonds."""
    dims = [int(d) for d in dims]
    total_dim, num_dims = 1, 1
    while num_dims < len(dims):
        total_dim *= dims[num_dims - 1]
        if total_dim == number_partitions:
            break
        elif total_dim > number_partitions:
            _log.error("Partitions and product of dimensions must match")
        num_dims += 1
    sub = dims[:num_dims]
    size = 1
    for d in sub:
        size *= d
    if not 1 <= partition <= size:
        raise IndexError("partition %d out of range for dims %r" % (partition, sub))
    rem = partition - 1
    out = []
    for d in sub:
        out.append(rem % d + 1)
        rem //= d
    return tuple(out)


def partition_network_on_virtual_bonds(network: TensorNetworkCircuit,
                                       number_partitions: int, partition: int):
    """``src/layer2/slicing.jl:37-57``: virtual bonds in edge insertion order,
    bond dims read from the *data* of the src node (``load_tensor_data``),
    stable sort by dim, then ``multi_index_partition``."""
    virtual_bonds = [(k, e) for k, e in network.edges.items() if e.virtual]
    bond_dims: List[int] = []
    for edge_index, edge in virtual_bonds:
        node_data = network.load_tensor_data(edge.src)
        shape = node_data.shape
        d = 1
        for pos, x in enumerate(network.nodes[edge.src].indices):
            if x == edge_index:
                d *= int(shape[pos])
        bond_dims.append(d)
    order = sorted(range(len(bond_dims)), key=lambda i: bond_dims[i])  # stable, like sortperm
    labels = [virtual_bonds[i][0] for i in order]
    bond_dims = [bond_dims[i] for i in order]
    assert len(bond_dims) > 0, "There must be some virtual bonds, try turning on decompose"
    ci = multi_index_partition(tuple(bond_dims), number_partitions, partition)
    return labels[:len(ci)], ci


def replace_with_view(network: TensorNetworkCircuit, node_label: str, bond_label: str,
                      bond_range: Sequence[int]) -> str:
    """``src/layer2/slicing.jl:67-93``.  The replacement node keeps the old
    ``dims`` (stale, App. D.2); the backend copies the sub-block."""
    node = network.nodes[node_label]
    bond_idx = node.indices.index(bond_label) + 1
    view_node = new_label(network, "node")
    network.nodes[view_node] = Node(node.indices, node.dims, view_node)
    network.node_layers[view_node] = network.node_layers[node_label]
    for edge_index in node.indices:
        edge = network.edges[edge_index]
        if edge.src == node_label:
            edge.src = view_node
        else:
            edge.dst = view_node
    network.view_tensor(view_node, node_label, bond_idx, bond_range)
    network.delete_tensor(node_label)
    del network.nodes[node_label]
    del network.node_layers[node_label]
    return view_node


def slice_tensor_network(network: TensorNetworkCircuit, bond_labels: Sequence[str],
                         bond_values: Sequence[int]) -> None:
    """``src/layer2/slicing.jl:100-110``: for every sliced bond replace both
    end nodes by unit-range views at the bond value."""
    for bond_label, bond_value in zip(bond_labels, tuple(bond_values)):
        node_1 = network.edges[bond_label].src
        node_2 = network.edges[bond_label].dst
        replace_with_view(network, node_1, bond_label, range(bond_value, bond_value + 1))
        replace_with_view(network, node_2, bond_label, range(bond_value, bond_value + 1))
