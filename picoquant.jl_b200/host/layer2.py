"""Contraction drivers (host bookkeeping that walks a plan and issues backend
calls).  Restates ``src/layer2.jl``: ``sort_indices`` (:209-216),
``create_ncon_indices`` (:226-240), ``contract_pair!`` (:329-405),
``contract_network!`` for edge plans (:250-284) and node-pair plans
(:294-321), ``full_wavefunction_contraction!`` (:132-195),
``random_contraction_plan`` (:119-122), and the SVD-based bond operations
``decompose_tensor!`` (:487-559), ``compress_bond!`` (:450-472) and
``compress_tensor_chain!`` (:423-440), and the MPS drivers
``contract_mps_tensor_network_circuit!`` (:571-624) and
``calculate_mps_amplitudes!`` (:633-643), and the bond-merging in-order driver
``merge_common_bonds!`` (:25-69) / ``inorder_contraction!`` (:76-112).

All of this is integer / label work and must be bit-exact with the reference:
the emitted backend call stream (labels, ncon index lists, permutation and
reshape arguments) is compared verbatim against hand-derived ``.tl`` streams in
``tests/test_plan_stream.py``.
"""
from __future__ import annotations

import random
from typing import Dict, List, Optional, Sequence, Union

from .backends import record_compute_costs
from .layer3 import Edge, Node, TensorNetworkCircuit, inneighbours, new_label, _label_number


def sort_indices(A: Node, B: Node):
    """``src/layer2.jl:209-216``.  Julia's set operations on arrays keep
    first-appearance order: common = A's shared indices in A order;
    uncommon = A-open (A order) followed by B-open (B order)."""
    b_set = set(B.indices)
    common, seen = [], set()
    for x in A.indices:
        if x in b_set and x not in seen:
            common.append(x)
            seen.add(x)
    common_set = set(common)
    uncommon, seen = [], set()
    for x in A.indices:
        if x not in common_set and x not in seen:
            uncommon.append(x)
            seen.add(x)
    for x in B.indices:
        if x not in common_set and x not in seen:
            uncommon.append(x)
            seen.add(x)
    return common, uncommon


def create_ncon_indices(A: Node, B: Node, common_indices: Sequence[str],
                        uncommon_indices: Sequence[str]):
    """``src/layer2.jl:226-240``: contracted index k of ``common`` -> +k,
    open index k of ``uncommon`` -> -k (both 1-based)."""
    index_map: Dict[str, int] = {}
    for k, x in enumerate(common_indices, start=1):
        index_map[x] = k
    for k, x in enumerate(uncommon_indices, start=1):
        index_map[x] = -k
    return [index_map[x] for x in A.indices], [index_map[x] for x in B.indices]


def contract_pair(network: TensorNetworkCircuit, A_label: str,
                  B_label: Optional[str] = None) -> Optional[str]:
    """``contract_pair!``: with one label it is the edge form
    (``src/layer2.jl:329-337``, silently skipping edges that are already gone);
    with two labels it is the node form (``src/layer2.jl:346-405``)."""
    if B_label is None:
        edge = A_label
        if edge in network.edges:
            e = network.edges[edge]
            return contract_pair(network, e.src, e.dst)
        return None

    if A_label == B_label:
        return None

    A = network.nodes[A_label]
    B = network.nodes[B_label]
    common_indices, remaining_indices = sort_indices(A, B)
    A_ncon_indices, B_ncon_indices = create_ncon_indices(A, B, common_indices,
                                                         remaining_indices)

    # costs use the graph's dims (stale after slicing)
    dims_map: Dict[str, int] = {}
    for ind, d in zip(list(A.indices) + list(B.indices), list(A.dims) + list(B.dims)):
        dims_map[ind] = d
    contracted_dims = [dims_map[ind] for ind in common_indices]
    C_dims = [dims_map[ind] for ind in remaining_indices]
    record_compute_costs(network.backend, C_dims, contracted_dims)

    C_label = new_label(network, "node")
    network.nodes[C_label] = Node(remaining_indices, C_dims, C_label)

    for index in common_indices:
        del network.edges[index]

    for index in remaining_indices:
        e = network.edges[index]
        if e.src in (A_label, B_label):
            e.src = C_label
        elif e.dst in (A_label, B_label):
            e.dst = C_label

    del network.nodes[A_label]
    del network.nodes[B_label]

    network.contract_tensors(A_label, A_ncon_indices, B_label, B_ncon_indices, C_label)
    return C_label


def random_contraction_plan(network: TensorNetworkCircuit, rng: Optional[random.Random] = None):
    """``src/layer2.jl:119-122``: closed edges in random order.  (The
    reference's ``shuffle`` draws from Julia's global RNG; the stream cannot be
    matched here, so a ``random.Random`` may be supplied for determinism.)"""
    closed = [k for k, v in network.edges.items() if v.src is not None and v.dst is not None]
    (rng or random).shuffle(closed)
    return closed


def _finish(network: TensorNetworkCircuit, output_shape) -> str:
    # common tail of both contract_network! methods (layer2.jl:266-283 / 303-320)
    output_tensor = "node_%d" % network.counters["node"]
    node = network.nodes[output_tensor]
    if len(node.indices) != 0:
        order = [node.indices.index(ind) + 1 for ind in network.output_qubits]
        node.indices[:] = network.output_qubits[:]
        network.permute_tensor(output_tensor, order)
        if output_shape == "vector":
            network.reshape_tensor(output_tensor, [list(range(1, len(node.indices) + 1))])
        elif output_shape != "":
            network.reshape_tensor(output_tensor, output_shape)
    network.save_output(output_tensor)
    return output_tensor


def contract_network(network: TensorNetworkCircuit,
                     plan: Union[Sequence[str], Sequence[Sequence[str]]],
                     output_shape: Union[str, List[List[int]]] = "") -> str:
    """``contract_network!``.  ``plan`` is either a list of edge labels
    (``src/layer2.jl:250-284``; leftover disjoint pieces are then contracted
    first-two-in-insertion-order) or a list of ``[A, B]`` node-label pairs
    (``src/layer2.jl:294-321``)."""
    plan = list(plan)
    if plan and not isinstance(plan[0], str):
        for A, B in plan:
            contract_pair(network, A, B)
        return _finish(network, output_shape)

    for edge in plan:
        contract_pair(network, edge)
    while len(network.nodes) > 1:
        it = iter(network.nodes)
        n1 = next(it)
        n2 = next(it)
        contract_pair(network, n1, n2)
    return _finish(network, output_shape)


def _layer_nodes(network: TensorNetworkCircuit) -> Dict[int, List[str]]:
    """``src/layer2.jl:149-156`` builds layer -> nodes by iterating an unordered
    ``Dict`` (hash order in Julia).  The mirror fixes the
    within-layer order to ascending node number, which for an SVD-split gate is
    (first half, second half)."""
    layer_nodes: Dict[int, List[str]] = {}
    for k in sorted(network.node_layers, key=_label_number):
        layer_nodes.setdefault(network.node_layers[k], []).append(k)
    return layer_nodes


def full_wavefunction_contraction(network: TensorNetworkCircuit,
                                  output_shape: Union[str, List[List[int]]] = "") -> str:
    """``full_wavefunction_contraction!`` (``src/layer2.jl:132-195``): outer
    product of the input caps, then one ``contract_pair`` per gate layer in
    circuit order, output caps (layer -1) last, final permute to
    ``output_qubits[qubit_ordering]`` order, optional reshape, ``save_output``.
    Returns the *label* of the final tensor (App. D.4)."""
    input_nodes = [network.edges[e].src for e in network.input_qubits]
    if any(n is None for n in input_nodes):
        raise RuntimeError("Please create nodes for each input before using wavefunction contraction")

    wf = input_nodes[0]
    for wfi in input_nodes[1:]:
        wf = contract_pair(network, wf, wfi)

    layer_nodes = _layer_nodes(network)
    gate_layers = [k for k in sorted(layer_nodes) if k > 0]
    if -1 in layer_nodes:
        gate_layers.append(-1)
    for layer in gate_layers:
        nodes = layer_nodes[layer]
        n = nodes[0]
        for n2 in nodes[1:]:
            n = contract_pair(network, n, n2)
        wf = contract_pair(network, wf, n)

    node = network.nodes[wf]
    if len(node.indices) != 0:
        output_qubits = [network.output_qubits[q - 1] for q in network.qubit_ordering]
        order = [node.indices.index(ind) + 1 for ind in output_qubits]
        network.permute_tensor(wf, order)
        if output_shape == "vector":
            network.reshape_tensor(wf, [list(range(1, len(node.indices) + 1))])
        elif output_shape != "":
            network.reshape_tensor(wf, output_shape)

    network.save_output(wf)
    return wf


# ---------------------------------------------------------------------------
# compression of a tensor network (src/layer2.jl:407-559)
# ---------------------------------------------------------------------------
def decompose_tensor(network: TensorNetworkCircuit, node_label: str,
                     left_indices: Sequence[str], right_indices: Sequence[str], *,
                     threshold: float = 1e-13, max_rank: int = 0,
                     left_label: Optional[str] = None, right_label: Optional[str] = None):
    """Network-level ``decompose_tensor!`` (``src/layer2.jl:487-559``): splits
    node ``node_label`` into two nodes joined by a new virtual edge; the index
    lists name the edges that go to either side.  The backend performs the SVD
    and returns the bond dimension chi (0 from a DSL backend: unknown until run
    time, in which case the upper bound min(left, right[, max_rank]) is used
    for the graph-side dims).  Returns ``(B_label, C_label)``."""
    node = network.nodes[node_label]
    index_map = {v: k for k, v in enumerate(node.indices, start=1)}
    left_positions = [index_map[x] for x in left_indices]
    right_positions = [index_map[x] for x in right_indices]

    B_label = new_label(network, "node") if left_label is None else left_label
    C_label = new_label(network, "node") if right_label is None else right_label

    chi = network.decompose_tensor(node_label, left_positions, right_positions,
                                   threshold=threshold, max_rank=max_rank,
                                   left_label=B_label, right_label=C_label)

    B_dims = [node.dims[i - 1] for i in left_positions]
    C_dims = [node.dims[i - 1] for i in right_positions]
    if chi > 0:
        virtual_dim = chi
    else:
        left_dim, right_dim = 1, 1
        for d in B_dims:
            left_dim *= d
        for d in C_dims:
            right_dim *= d
        virtual_dim = min(left_dim, right_dim)
        if max_rank > 0:
            virtual_dim = min(virtual_dim, max_rank)
    B_dims.append(virtual_dim)
    C_dims.insert(0, virtual_dim)

    index_label = new_label(network, "index")
    B_node = Node(list(left_indices) + [index_label], B_dims, B_label)
    C_node = Node([index_label] + list(right_indices), C_dims, C_label)
    # the reference deletes the original node last (layer2.jl:555-556); when a label is
    # re-used (compress_bond!) the new node must survive that, so drop the old entry first
    network.nodes.pop(node_label, None)
    network.nodes[B_label] = B_node
    network.nodes[C_label] = C_node

    for index in left_indices:
        e = network.edges[index]
        if e.src == node_label:
            e.src = B_label
        elif e.dst == node_label:
            e.dst = B_label
    for index in right_indices:
        e = network.edges[index]
        if e.src == node_label:
            e.src = C_label
        elif e.dst == node_label:
            e.dst = C_label
    network.edges[index_label] = Edge(B_label, C_label, None, True)

    if node_label not in (B_label, C_label):
        network.delete_tensor(node_label)
    return B_label, C_label


def compress_bond(network: TensorNetworkCircuit, node_1: str, node_2: str, *,
                  threshold: float = 1e-13, max_rank: int = 0):
    """``compress_bond!`` (``src/layer2.jl:450-472``): contract the two nodes,
    then split the result again along the same bipartition, discarding singular
    values below the threshold / beyond ``max_rank``."""
    left_node = network.nodes[node_1]
    right_node = network.nodes[node_2]
    right_set, left_set = set(right_node.indices), set(left_node.indices)
    left_indices = [x for x in left_node.indices if x not in right_set]
    right_indices = [x for x in right_node.indices if x not in left_set]
    combined = contract_pair(network, node_1, node_2)
    return decompose_tensor(network, combined, left_indices, right_indices,
                            left_label=node_1, right_label=node_2,
                            threshold=threshold, max_rank=max_rank)


def compress_tensor_chain(network: TensorNetworkCircuit, nodes: Sequence[str], *,
                          threshold: float = 1e-13, max_rank: int = 0) -> None:
    """``compress_tensor_chain!`` (``src/layer2.jl:423-440``): forward then
    backward sweep of ``compress_bond!`` over consecutive nodes."""
    nodes = list(nodes)
    for i in range(len(nodes) - 1):
        compress_bond(network, nodes[i], nodes[i + 1], threshold=threshold, max_rank=max_rank)
    for i in range(len(nodes) - 2, -1, -1):
        compress_bond(network, nodes[i], nodes[i + 1], threshold=threshold, max_rank=max_rank)


# ---------------------------------------------------------------------------
# MPS contraction of a circuit (src/layer2.jl:563-643)
# ---------------------------------------------------------------------------
def contract_mps_tensor_network_circuit(network: TensorNetworkCircuit, *, max_bond: int = 2,
                                        threshold: float = 1e-13, max_rank: int = 0,
                                        pbc: bool = False) -> List[str]:
    """``contract_mps_tensor_network_circuit!`` (``src/layer2.jl:571-624``): the input
    caps are the MPS sites; gate layers are absorbed in circuit order (``contract_pair!``
    of the site with the gate half acting on it) and, after every layer, the bonds between
    the touched neighbouring sites are compressed (``compress_bond!``: contract + SVD).
    Gates must act on neighbouring qubits (``decompose=true`` networks).  Returns the site
    labels, each also saved as an output under its own name."""
    mps_nodes = [network.edges[x].src for x in network.input_qubits]
    assert all(x is not None for x in mps_nodes), "Input qubit values must be set"
    layer_nodes = _layer_nodes(network)
    gate_layers = [k for k in sorted(layer_nodes) if k > 0]
    if -1 in layer_nodes:
        gate_layers.append(-1)
    for gate_layer in gate_layers:
        updated = []
        for node in layer_nodes[gate_layer]:
            input_node = inneighbours(network, node)[0]
            assert input_node in mps_nodes, "%s not in mps nodes" % input_node
            idx = mps_nodes.index(input_node)
            mps_nodes[idx] = contract_pair(network, input_node, node)
            updated.append(idx)
        updated.sort()
        for i in range(len(updated) - 1):
            distance = updated[i + 1] - updated[i]
            if pbc:
                distance = min(distance, updated[i] + len(mps_nodes) - updated[i + 1])
            assert distance == 1, "Gates between non-neighboring qubits"
            compress_bond(network, mps_nodes[updated[i]], mps_nodes[updated[i + 1]],
                          threshold=threshold, max_rank=max_rank)
    for node in mps_nodes:
        network.save_output(node, node)
    return mps_nodes


def contract_tensor_network_circuit_with_compression(network: TensorNetworkCircuit, *,
                                                     max_bond: int = 2, threshold: float = 1e-13,
                                                     max_rank: int = 0) -> List[str]:
    """``contract_tensor_network_circuit_with_compression!`` (``src/layer2.jl:657-702``): like
    the MPS driver but without the neighbour requirement -- after every layer the bonds
    between consecutively touched sites are compressed in the order the gates were met."""
    network_nodes = [network.edges[x].src for x in network.input_qubits]
    assert all(x is not None for x in network_nodes), "Input qubit values must be set"
    layer_nodes = _layer_nodes(network)
    gate_layers = [k for k in sorted(layer_nodes) if k > 0]
    if -1 in layer_nodes:
        gate_layers.append(-1)
    for gate_layer in gate_layers:
        updated = []
        for node in layer_nodes[gate_layer]:
            input_node = inneighbours(network, node)[0]
            assert input_node in network_nodes, "%s not in network nodes" % input_node
            idx = network_nodes.index(input_node)
            network_nodes[idx] = contract_pair(network, input_node, node)
            updated.append(idx)
        for i in range(len(updated) - 1):
            compress_bond(network, network_nodes[updated[i]], network_nodes[updated[i + 1]],
                          threshold=threshold, max_rank=max_rank)
    for node in network_nodes:
        network.save_output(node, node)
    return network_nodes


def calculate_mps_amplitudes(network: TensorNetworkCircuit, mps_nodes: Sequence[str],
                             result: str = "result") -> None:
    """``calculate_mps_amplitudes!`` (``src/layer2.jl:633-643``): contract the chain left to
    right, permute by ``qubit_ordering``, flatten, save."""
    output_node = mps_nodes[0]
    for node in mps_nodes[1:]:
        output_node = contract_pair(network, output_node, node)
    network.permute_tensor(output_node, list(network.qubit_ordering))
    network.reshape_tensor(output_node, [list(range(1, len(mps_nodes) + 1))])
    network.save_output(output_node, result)


# ---------------------------------------------------------------------------
# in-order contraction with bond merging (src/layer2.jl:25-112)
# ---------------------------------------------------------------------------
def merge_common_bonds(network: TensorNetworkCircuit, a_label: str, b_label: str) -> None:
    """``merge_common_bonds!`` (``src/layer2.jl:25-69``): when two nodes share more than one
    index, both tensors are permuted to [remaining..., common...] and the common axes are
    fused into one (``permute_tensor`` + ``reshape_tensor`` on the backend); the shared edges
    are replaced by a single new virtual edge."""
    a = network.nodes[a_label]
    b = network.nodes[b_label]
    b_set, a_set = set(b.indices), set(a.indices)
    common_edges = [x for x in a.indices if x in b_set]
    if len(common_edges) <= 1:
        return
    a_common = [k for k, x in enumerate(a.indices, start=1) if x in b_set]
    a_remaining = [k for k, x in enumerate(a.indices, start=1) if x not in b_set]
    network.permute_tensor(a_label, a_remaining + a_common)
    b_common = [k for k, x in enumerate(b.indices, start=1) if x in a_set]
    b_remaining = [k for k, x in enumerate(b.indices, start=1) if x not in a_set]
    network.permute_tensor(b_label, b_remaining + b_common)

    new_edge = new_label(network, "index")
    indices_a = [a.indices[k - 1] for k in a_remaining] + [new_edge]
    indices_b = [b.indices[k - 1] for k in b_remaining] + [new_edge]
    dims_map_a = dict(zip(a.indices, a.dims))
    dims_map_b = dict(zip(b.indices, b.dims))
    merged = 1
    for k in a_common:
        merged *= dims_map_a[a.indices[k - 1]]
    dims_map_a[new_edge] = dims_map_b[new_edge] = merged
    network.nodes[a_label] = Node(indices_a, [dims_map_a[i] for i in indices_a], a_label)
    network.nodes[b_label] = Node(indices_b, [dims_map_b[i] for i in indices_b], b_label)

    l, m = len(a_remaining), len(a_common)
    network.reshape_tensor(a_label, [[x] for x in range(1, l + 1)] + [list(range(l + 1, l + m + 1))])
    l, m = len(b_remaining), len(b_common)
    network.reshape_tensor(b_label, [[x] for x in range(1, l + 1)] + [list(range(l + 1, l + m + 1))])

    network.edges[new_edge] = Edge(a_label, b_label, None, True)
    for e in common_edges:
        del network.edges[e]


def inorder_contraction(network: TensorNetworkCircuit) -> None:
    """``inorder_contraction!`` (``src/layer2.jl:76-112``): layer by layer, every gate node
    absorbs its (unique) in-neighbours; afterwards multiple bonds between the new nodes of
    a layer are merged.  ``y > x`` compares the labels as strings, like Julia Symbols."""
    layer_nodes = _layer_nodes(network)
    gate_layers = [k for k in sorted(layer_nodes) if k > 0]
    if -1 in layer_nodes:
        gate_layers.append(-1)
    for layer in gate_layers:
        new_nodes = []
        for node in layer_nodes[layer]:
            in_nodes, seen = [], set()
            for n in inneighbours(network, node):
                if n not in seen:
                    seen.add(n)
                    in_nodes.append(n)
            for in_node in in_nodes:
                node = contract_pair(network, in_node, node)
            new_nodes.append(node)
        # Iterators.product(new_nodes, new_nodes): the first factor varies fastest
        for y in new_nodes:
            for x in new_nodes:
                if y > x:
                    merge_common_bonds(network, x, y)
