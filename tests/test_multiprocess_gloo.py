"""World-size-2 check of the sharded slice loop on CPU (gloo): every rank takes
its contiguous block of partitions, accumulates locally and one all-reduce sums
the partials -- the host-side logic of the multi-GPU path, with the oracle
backend standing in for the device."""
import os
import socket
import sys

import numpy as np
import pytest

import picoquant_jl_b200  # noqa: F401
from picoquant_jl_b200.host.sliced import partitions_of_rank

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_blocks_cover_all_slices():
    for P in (1, 2, 4, 8, 64):
        for world in (1, 2, 3, 4, 8):
            got = [p for r in range(world) for p in partitions_of_rank(P, r, world)]
            assert got == list(range(1, P + 1))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import picoquant_jl_b200  # noqa: F401
    from oracle.interactive import OracleBackend, execute_dsl
    from picoquant_jl_b200.host import create_qft_circuit, create_simple_preparation_circuit
    from picoquant_jl_b200.host.backends import TensorStore
    from picoquant_jl_b200.host.sliced import partitions_of_rank, record_sliced_contraction

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank,
                            world_size=world)
    n, P = 4, 8
    circ = create_simple_preparation_circuit(n, 2, 5).compose(create_qft_circuit(n))
    rec = record_sliced_contraction(circ, P, 1, output_shape="vector")
    partial = np.zeros(2 ** n, dtype=np.complex128)
    for p in partitions_of_rank(P, rank, world):
        out = TensorStore()
        execute_dsl(rec.text_for(p), rec.store, np.complex128, output_store=out)
        partial += out.read("result")
    t = torch.from_numpy(np.stack([partial.real, partial.imag]))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    total = t[0].numpy() + 1j * t[1].numpy()
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), total)
    dist.destroy_process_group()


def test_two_rank_sliced_sum_matches_unsliced(tmp_path):
    import torch.multiprocessing as mp
    from helpers import rel_l2, statevector
    from oracle.interactive import OracleBackend
    from picoquant_jl_b200.host import create_qft_circuit, create_simple_preparation_circuit

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    circ = create_simple_preparation_circuit(4, 2, 5).compose(create_qft_circuit(4))
    ref = statevector(circ, OracleBackend(np.complex128), decompose=True)
    for r in range(2):
        got = np.load(tmp_path / ("rank%d.npy" % r))
        assert rel_l2(got, ref) < 1e-13
