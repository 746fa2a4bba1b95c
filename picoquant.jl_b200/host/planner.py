"""Deterministic contraction planners for the harness.

The reference's planners (netcon / bgreedy / QuickBB / KaHyPar,
``src/layer2/*_contraction.jl``) are outside the hot path and either limited to
<= 36 nodes or unusable at RQC scale.  Plans
are *inputs* to the hot path, so the harness brings its own planner.  Its
output has exactly the shape the reference consumes: a list of ``[A, B]``
node-label pairs for ``contract_network!(network, plan)``
(``src/layer2.jl:294-321``), with intermediate labels predicted the way
``convert_tree_to_plan`` does (``src/layer2/netcon_contraction.jl:97-123``):
every ``contract_pair!`` creates ``node_{counter+1}``.

Two strategies, both deterministic for a given seed:

* ``greedy_plan``      -- size-difference greedy with seeded Boltzmann noise,
  best of ``trials`` by total MAC count subject to a max-intermediate cap;
* ``sweep_plan``       -- geometry-aware plan for circuits on a grid of qubits:
  collapse every qubit's world-line into one site tensor, then absorb the
  sites along a snake order (exact boundary contraction).  This is what makes
  the 7x7 depth-24 amplitude tractable.
"""
from __future__ import annotations

import heapq
import math
import random
from typing import Dict, List, Optional, Sequence, Tuple

from .layer3 import TensorNetworkCircuit, _label_number


class _Graph:
    """Light copy of the network topology with *effective* index extents."""

    def __init__(self, network: TensorNetworkCircuit,
                 sliced_bonds: Sequence[str] = ()) -> None:
        sliced = set(sliced_bonds)
        self.dim: Dict[str, int] = {}
        self.nodes: Dict[str, List[str]] = {}
        for label, node in network.nodes.items():
            self.nodes[label] = list(node.indices)
            for ind, d in zip(node.indices, node.dims):
                self.dim[ind] = 1 if ind in sliced else int(d)
        self.next_id = network.counters["node"]

    def size(self, indices: Sequence[str]) -> int:
        s = 1
        for i in indices:
            s *= self.dim[i]
        return s


def _contract_indices(a: Sequence[str], b: Sequence[str]):
    bset = set(b)
    common = [x for x in a if x in bset]
    cset = set(common)
    rest = [x for x in a if x not in cset] + [x for x in b if x not in cset]
    return common, rest


def plan_cost(network: TensorNetworkCircuit, plan: Sequence[Sequence[str]],
              sliced_bonds: Sequence[str] = ()) -> Dict[str, float]:
    """Replays ``plan`` symbolically; returns total complex MACs, the largest
    intermediate (elements) and the per-step (M, N, K) list."""
    g = _Graph(network, sliced_bonds)
    nodes = dict(g.nodes)
    nid = g.next_id
    macs, biggest, steps = 0, 0, []
    for a, b in plan:
        ia, ib = nodes.pop(a), nodes.pop(b)
        common, rest = _contract_indices(ia, ib)
        K = g.size(common)
        cs = set(common)
        M = g.size([x for x in ia if x not in cs])
        N = g.size([x for x in ib if x not in cs])
        macs += M * N * K
        biggest = max(biggest, M * N)
        steps.append((M, N, K))
        nid += 1
        nodes["node_%d" % nid] = rest
    return {"macs": macs, "max_size": biggest, "steps": steps, "remaining": len(nodes)}


def greedy_plan(network: TensorNetworkCircuit, *, sliced_bonds: Sequence[str] = (),
                alpha: float = 1.0, temperature: float = 0.0, trials: int = 1,
                seed: int = 0, max_size: Optional[int] = None) -> List[List[str]]:
    """Greedy pairwise plan: repeatedly contract the connected pair minimising
    ``size(C) - alpha * (size(A) + size(B))`` (plus Gumbel noise scaled by
    ``temperature`` in log-space).  Disconnected leftovers are merged smallest
    first.  Returns the best of ``trials`` runs (trial 0 is noise-free)."""
    best, best_key = None, None
    for t in range(max(1, trials)):
        rng = random.Random(seed * 1000003 + t)
        plan = _greedy_once(network, sliced_bonds, alpha, temperature if t else 0.0, rng)
        c = plan_cost(network, plan, sliced_bonds)
        over = 0 if max_size is None or c["max_size"] <= max_size else 1
        key = (over, c["macs"], c["max_size"])
        if best_key is None or key < best_key:
            best, best_key = plan, key
    return best


def _greedy_once(network, sliced_bonds, alpha, temperature, rng) -> List[List[str]]:
    g = _Graph(network, sliced_bonds)
    nodes: Dict[str, List[str]] = dict(g.nodes)
    order: Dict[str, int] = {k: _label_number(k) for k in nodes}
    owners: Dict[str, List[str]] = {}
    for label, inds in nodes.items():
        for i in inds:
            owners.setdefault(i, []).append(label)
    nid = g.next_id
    heap: List[Tuple[float, int, int, str, str]] = []

    def score(a: str, b: str) -> float:
        ia, ib = nodes[a], nodes[b]
        _, rest = _contract_indices(ia, ib)
        s = g.size(rest) - alpha * (g.size(ia) + g.size(ib))
        if temperature > 0.0:
            # multiplicative noise on the magnitude keeps the sign structure
            gumbel = -math.log(-math.log(rng.random() or 1e-300))
            s = s - temperature * gumbel * max(1.0, abs(s))
        return s

    def push(a: str, b: str) -> None:
        if order[a] > order[b]:
            a, b = b, a
        heapq.heappush(heap, (score(a, b), order[a], order[b], a, b))

    seen = set()
    for i, own in owners.items():
        if len(own) == 2 and own[0] != own[1]:
            key = (own[0], own[1]) if order[own[0]] < order[own[1]] else (own[1], own[0])
            if key not in seen:
                seen.add(key)
                push(*key)

    plan: List[List[str]] = []
    while heap:
        _, _, _, a, b = heapq.heappop(heap)
        if a not in nodes or b not in nodes:
            continue
        ia, ib = nodes.pop(a), nodes.pop(b)
        _, rest = _contract_indices(ia, ib)
        nid += 1
        c = "node_%d" % nid
        nodes[c] = rest
        order[c] = nid
        plan.append([a, b])
        neigh = []
        for i in rest:
            own = owners[i]
            for k in range(len(own)):
                if own[k] == a or own[k] == b:
                    own[k] = c
            for o in own:
                if o != c and o not in neigh:
                    neigh.append(o)
        for i in ia:
            if i not in rest:
                owners.pop(i, None)
        for o in neigh:
            push(c, o)

    # disconnected pieces (outer products): smallest first, deterministic
    while len(nodes) > 1:
        a, b = sorted(nodes, key=lambda k: (g.size(nodes[k]), order[k]))[:2]
        if order[a] > order[b]:
            a, b = b, a
        ia, ib = nodes.pop(a), nodes.pop(b)
        nid += 1
        c = "node_%d" % nid
        nodes[c] = ia + ib
        order[c] = nid
        plan.append([a, b])
    return plan


def sweep_plan(network: TensorNetworkCircuit, rows: int, cols: int, *,
               sliced_bonds: Sequence[str] = (), two_sided: bool = True) -> List[List[str]]:
    """Geometry-aware plan for a ``rows x cols`` grid circuit (qubit (i, j) is
    circuit qubit ``i + (j-1)*rows``, 1-based, as in ``create_RQC``).

    1. every node is assigned to a grid site through the ``qubit`` tag of its
       non-virtual edges (``Edge.qubit``, ``src/layer3.jl:43-53``); each site's
       world-line is contracted in time order into one site tensor;
    2. sites are absorbed column by column in snake order into a boundary
       tensor -- from the left and (``two_sided``) from the right, meeting in
       the middle with one final contraction.
    """
    g = _Graph(network, sliced_bonds)
    nodes: Dict[str, List[str]] = dict(g.nodes)
    nid = g.next_id
    plan: List[List[str]] = []

    site_of: Dict[str, int] = {}
    for label, inds in nodes.items():
        qs = [network.edges[i].qubit for i in inds if network.edges[i].qubit is not None]
        if not qs:
            raise ValueError("node %s has no qubit-tagged edge" % label)
        if len(set(qs)) != 1:
            raise ValueError("sweep_plan needs decompose=True (node %s spans qubits %r)"
                             % (label, sorted(set(qs))))
        site_of[label] = qs[0]

    def contract(a: str, b: str) -> str:
        nonlocal nid
        ia, ib = nodes.pop(a), nodes.pop(b)
        _, rest = _contract_indices(ia, ib)
        nid += 1
        c = "node_%d" % nid
        nodes[c] = rest
        plan.append([a, b])
        return c

    # 1. world-lines: circuit (node-number) order, caps included
    site_tensor: Dict[int, str] = {}
    per_site: Dict[int, List[str]] = {}
    for label in sorted(nodes, key=_label_number):
        per_site.setdefault(site_of[label], []).append(label)

    def time_key(label: str):
        layer = network.node_layers.get(label, 0)
        # input caps (layer 0) first, gates by layer, output caps (layer -1) last
        return (1 if layer == -1 else 0, layer, _label_number(label))

    for q, labels in per_site.items():
        labels = sorted(labels, key=time_key)
        cur = labels[0]
        for nxt in labels[1:]:
            cur = contract(cur, nxt)
        site_tensor[q] = cur

    def qubit(i: int, j: int) -> int:
        return i + (j - 1) * rows

    def snake(col_range) -> List[int]:
        out = []
        for n, j in enumerate(col_range):
            rng = range(1, rows + 1) if n % 2 == 0 else range(rows, 0, -1)
            out.extend(qubit(i, j) for i in rng if qubit(i, j) in site_tensor)
        return out

    if two_sided and cols >= 2:
        mid = (cols + 1) // 2
        left = snake(range(1, mid + 1))
        right = snake(range(cols, mid, -1))
    else:
        left, right = snake(range(1, cols + 1)), []

    def absorb(seq: List[int]) -> Optional[str]:
        if not seq:
            return None
        cur = site_tensor[seq[0]]
        for q in seq[1:]:
            cur = contract(cur, site_tensor[q])
        return cur

    l, r = absorb(left), absorb(right)
    if l is not None and r is not None:
        contract(l, r)
    return plan


# ---------------------------------------------------------------------------
# the reference's own planners, restated (src/layer2/bgreedy_contraction.jl,
# src/layer2/netcon_contraction.jl).  Plans are inputs to the hot path; these exist so
# that scripts written against PicoQuant's planner API keep working.
# ---------------------------------------------------------------------------
def contraction_cost(A, B, alpha: float):
    """``contraction_cost`` (``bgreedy_contraction.jl:73-91``): Gray-Kourtis heuristic
    ``|C| - alpha (|A| + |B|)``, the result's indices / dims, and [time, space] costs."""
    from .layer2 import sort_indices
    _, c_indices = sort_indices(A, B)
    cset = set(c_indices)
    a_open = [d for i, d in zip(A.indices, A.dims) if i in cset]
    b_open = [d for i, d in zip(B.indices, B.dims) if i in cset]
    c_dims = a_open + b_open
    size_c = math.prod(c_dims)
    size_a = math.prod(A.dims)
    size_b = math.prod(B.dims)
    costs = [math.sqrt(size_a * size_b * size_c), size_c]
    return size_c - alpha * (size_a + size_b), c_indices, c_dims, costs


def bgreedy(network: TensorNetworkCircuit, alpha: float = 1.0, tau: float = 1.0,
            N: Optional[int] = None, rng: Optional[random.Random] = None):
    """``bgreedy`` (``bgreedy_contraction.jl:17-65``): contract a random pair drawn from
    the Boltzmann distribution ``exp(-cost / tau)`` over ALL pairs until one tensor is left.
    With ``N`` the best of N samples by time cost is returned as ``(plan, time_cost)``;
    without it one sample as ``(plan, time_cost, space_cost)``.  (StatsBase's sampler and
    Julia's RNG stream cannot be matched; pass ``rng`` for reproducibility.)"""
    from .layer3 import Node
    rng = rng or random.Random()

    def once():
        tensors = dict(network.nodes)
        plan: List[List[str]] = []
        count = network.counters["node"]
        time_cost, space_cost = 0.0, 0.0
        while len(tensors) > 1:
            items = list(tensors.items())
            pairs, energy, cinds, cdims, costs = [], [], [], [], []
            for i in range(len(items)):
                for j in range(i + 1, len(items)):
                    e, ci, cd, co = contraction_cost(items[i][1], items[j][1], alpha)
                    pairs.append((items[i][0], items[j][0]))
                    energy.append(e)
                    cinds.append(ci)
                    cdims.append(cd)
                    costs.append(co)
            emin = min(energy)     # shift: same distribution, no overflow in exp
            weights = [math.exp(-(e - emin) / tau) for e in energy]
            k = rng.choices(range(len(pairs)), weights=weights)[0]
            a, b = pairs[k]
            count += 1
            c = "node_%d" % count
            del tensors[a]
            del tensors[b]
            tensors[c] = Node(cinds[k], cdims[k], "intermediate_tensor")
            plan.append([a, b])
            time_cost += costs[k][0]
            space_cost = max(space_cost, costs[k][1])
        return plan, time_cost, space_cost

    if N is None:
        return once()
    best_plan, best_cost = [], math.inf
    for _ in range(N):
        plan, t, _ = once()
        if t < best_cost:
            best_plan, best_cost = plan, t
    return best_plan, best_cost


def bgreedy_contraction(network: TensorNetworkCircuit, alpha: float = 1.0, tau: float = 1.0,
                        N: int = 10, output_shape="", rng: Optional[random.Random] = None):
    """``bgreedy_contraction!`` (``bgreedy_contraction.jl:140-145``)."""
    from .layer2 import contract_network
    plan, _ = bgreedy(network, alpha, tau, N, rng)
    return contract_network(network, plan, output_shape)


def netcon(network: TensorNetworkCircuit) -> List[List[str]]:
    """``netcon`` (``netcon_contraction.jl:15-41``) returns the pair plan of a
    minimum-flop contraction tree.  The reference delegates the search to
    TensorOperations.jl's ``optimaltree`` (un-vendored); here the optimum over all binary
    trees is found by dynamic programming over subsets of tensors (cost of joining two
    groups = product of the dims of every index either group still exposes), which the
    reference's 36-node limit does not need but its tests (3-qubit circuits) fit easily:
    at most 16 tensors are accepted.  The tree is converted to a plan exactly like
    ``convert_tree_to_plan`` (``:97-123``): left subtree, right subtree, then the pair."""
    labels = list(network.nodes)
    n = len(labels)
    if n > 16:
        raise ValueError("netcon (subset DP) handles at most 16 tensors, got %d" % n)
    if n == 0:
        return []
    dims: Dict[str, int] = {}
    owners: Dict[str, int] = {}
    for t, lab in enumerate(labels):
        node = network.nodes[lab]
        for ind, d in zip(node.indices, node.dims):
            dims[ind] = int(d)
            owners[ind] = owners.get(ind, 0) | (1 << t)
    full = (1 << n) - 1

    def exposed(mask: int) -> int:
        # product of dims of indices with an owner inside and (an owner outside or open)
        p = 1
        for ind, own in owners.items():
            if own & mask and (own & ~mask & full or bin(own).count("1") == 1):
                p *= dims[ind]
        return p

    def join_cost(a: int, b: int) -> int:
        p = 1
        for ind, own in owners.items():
            if own & (a | b):
                inside_only = (own & ~(a | b) & full) == 0 and bin(own).count("1") > 1
                shared = bool(own & a) and bool(own & b)
                if shared or not inside_only:
                    p *= dims[ind]
        return p

    best: Dict[int, Tuple[int, object]] = {1 << t: (0, t) for t in range(n)}
    for size in range(2, n + 1):
        for mask in range(1, full + 1):
            if bin(mask).count("1") != size:
                continue
            low = mask & -mask
            sub = (mask - 1) & mask
            cand = None
            while sub:
                if sub & low and sub != mask:       # each split once
                    other = mask ^ sub
                    if sub in best and other in best:
                        c = best[sub][0] + best[other][0] + join_cost(sub, other)
                        if cand is None or c < cand[0]:
                            cand = (c, (best[sub][1], best[other][1]))
                sub = (sub - 1) & mask
            if cand is not None:
                best[mask] = cand
    tree = best[full][1]
    plan: List[List[str]] = []
    counter = [network.counters["node"]]

    def convert(t):
        if isinstance(t, int):
            return labels[t]
        a = convert(t[0])
        b = convert(t[1])
        plan.append([a, b])
        counter[0] += 1
        return "node_%d" % counter[0]

    convert(tree)
    return plan


def netcon_contraction(network: TensorNetworkCircuit, output_shape=""):
    """``netcon_contraction!`` (``netcon_contraction.jl:152-160``)."""
    from .layer2 import contract_network
    return contract_network(network, netcon(network), output_shape)
