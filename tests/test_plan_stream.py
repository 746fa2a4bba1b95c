"""Bit-exact plan / slice indexing: the command stream emitted by the host
mirror through the DSL backend must equal, byte for byte, the stream the
reference code emits (hand-derived from src/layer2.jl:132-405,
src/layer2/slicing.jl and src/backends/dsl.jl:113-214)."""
import logging

import numpy as np
import pytest

from picoquant_jl_b200.host import (DSLBackend, TensorNetworkCircuit, add_gate, add_input,
                                    full_wavefunction_contraction, gate_tensor,
                                    multi_index_partition, parse_dsl,
                                    partition_network_on_virtual_bonds, slice_tensor_network)


def test_nb3_ghz_stream():
    """nb/3.The-DSL-backend.ipynb cell 2: add_input first, then H, CX, CX."""
    dsl = DSLBackend(tensor_data="tensor_file.h5", output="output_file.h5")
    tn = TensorNetworkCircuit(3, dsl)
    add_input(tn, "000")
    add_gate(tn, gate_tensor("H"), [1])
    add_gate(tn, gate_tensor("CX"), [1, 2])
    add_gate(tn, gate_tensor("CX"), [2, 3])
    out = full_wavefunction_contraction(tn, "vector")
    assert out == "node_11"
    expected = """tensor node_1 node_1
tensor node_2 node_2
tensor node_3 node_3
tensor node_4 node_4
tensor node_5 node_5
tensor node_6 node_6
ncon node_7 node_1 -1 node_2 -2
del node_1
del node_2
ncon node_8 node_7 -1,-2 node_3 -3
del node_7
del node_3
ncon node_9 node_8 1,-1,-2 node_4 1,-3
del node_8
del node_4
ncon node_10 node_9 1,-1,2 node_5 2,1,-2,-3
del node_9
del node_5
ncon node_11 node_10 1,-1,2 node_6 2,1,-2,-3
del node_10
del node_6
permute node_11 1,2,3
reshape node_11 1,2,3
save node_11 output_file.h5 result
"""
    assert dsl.text() == expected


def test_sliced_decomposed_cx_stream():
    """Decomposed CX on 2 qubits, P=2, partition 2: view / del / ncon lines."""
    dsl = DSLBackend()
    tn = TensorNetworkCircuit(2, dsl)
    labels = add_gate(tn, gate_tensor("CX"), [1, 2], decompose=True)
    assert labels == ["node_1", "node_2"]
    add_input(tn, "00")
    bonds, values = partition_network_on_virtual_bonds(tn, 2, 2)
    assert bonds == ["index_5"] and values == (2,)
    slice_tensor_network(tn, bonds, values)
    # the replacement nodes keep the unsliced dims (slicing.jl:77, App. D.2)
    assert tn.nodes["node_5"].dims == [2, 2, 2]
    full_wavefunction_contraction(tn, "vector")
    expected = """tensor node_1 node_1
tensor node_2 node_2
tensor node_3 node_3
tensor node_4 node_4
view node_5 node_1 3 2
del node_1
view node_6 node_2 1 2
del node_2
ncon node_7 node_3 -1 node_4 -2
del node_3
del node_4
ncon node_8 node_5 -1,-2,1 node_6 1,-3,-4
del node_5
del node_6
ncon node_9 node_7 1,2 node_8 1,-1,2,-2
del node_7
del node_8
permute node_9 1,2
reshape node_9 1,2
save node_9 tensor_data.h5 result
"""
    assert dsl.text() == expected
    ops = parse_dsl(dsl.text())
    assert ops[4] == ("view", dict(v="node_5", t="node_1", axis=3, idx=[2]))
    assert ops[11][1]["a_idx"] == [-1, -2, 1]


def test_multi_index_partition_values():
    """slicing.jl:12-26 -- column-major unravel of the 1-based partition id."""
    assert [multi_index_partition((2, 2, 2), 4, p) for p in (1, 2, 3, 4)] == \
        [(1, 1), (2, 1), (1, 2), (2, 2)]
    assert multi_index_partition((2, 4, 2), 8, 3) == (1, 2)
    assert multi_index_partition((2, 4, 2), 8, 8) == (2, 4)
    assert multi_index_partition((2,), 2, 2) == (2,)
    dims = (2,) * 10
    for p in (1, 37, 64):
        v = multi_index_partition(dims, 64, p)
        assert len(v) == 6
        assert sum((b - 1) << i for i, b in enumerate(v)) == p - 1
    with pytest.raises(IndexError):
        multi_index_partition((2, 2), 2, 3)


def test_multi_index_partition_quirk_single_partition(caplog):
    """App. D.1: number_partitions=1 logs an error and indexes over all bonds."""
    with caplog.at_level(logging.ERROR):
        assert multi_index_partition((2, 2), 1, 1) == (1, 1)
    assert "must match" in caplog.text


def test_multi_index_partition_is_a_bijection_property():
    """Property (hypothesis): for bond dims whose prefix product hits P exactly, the P partition
    ids map one-to-one onto the multi-indices of that prefix, first bond fastest -- the
    size-independent form of slicing.jl:12-26 that the sharded slice loop relies on."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=200, deadline=None)
    @given(st.lists(st.integers(min_value=2, max_value=5), min_size=1, max_size=5),
           st.lists(st.integers(min_value=2, max_value=4), min_size=1, max_size=3))
    def check(prefix, tail):
        dims = tuple(prefix + tail)         # the reference never inspects the last dim: keep a tail
        P = int(np.prod(prefix))
        seen = set()
        for p in range(1, P + 1):
            v = multi_index_partition(dims, P, p)
            assert len(v) == len(prefix)
            assert all(1 <= b <= d for b, d in zip(v, prefix))
            flat, stride = 0, 1
            for b, d in zip(v, prefix):
                flat += (b - 1) * stride
                stride *= d
            assert flat == p - 1            # column-major unravel
            seen.add(v)
        assert len(seen) == P

    check()
