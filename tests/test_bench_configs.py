"""bench.py --workload ghz3 | qft10 (BASELINE configs 1-2): the control flow and the JSON
contract of the per-config arm, exercised on the CPU.

The device backend is replaced by a stand-in built on the oracle (tests may use the oracle;
the product never does), so this checks bench.py's own logic -- stream accounting, the
reference arm, the keys of the line -- not the kernels.  The real thing runs on the GPU box:
``python bench.py --workload qft10``."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from oracle.interactive import OracleBackend, execute_dsl  # noqa: E402
from picoquant_jl_b200.host.backends import TensorStore  # noqa: E402


class _FakeProgram:
    def __init__(self, backend, text):
        self.backend, self.text = backend, text
        self.macs = bench.stream_costs(text, backend.leaves, backend.dtype.itemsize)["macs"]
        self.launches, self.arena_bytes = 7, 4096

    def run(self, *a):
        out = TensorStore()
        execute_dsl(self.text, self.backend.leaves, self.backend.dtype, output_store=out)
        OracleBackend.save_tensor_data(self.backend, "result", out.read("result"))
        self.backend.launches += self.launches

    def close(self):
        pass


class _FakeDevice(OracleBackend):
    def __init__(self, dtype=np.complex64, device=0):
        super().__init__(dtype)
        self.dtype = np.dtype(dtype)
        self.leaves = TensorStore()
        self.launches = 0

    def save_tensor_data(self, label, data):
        self.leaves.write(label, data)
        super().save_tensor_data(label, data)

    def compile_program(self, text):
        return _FakeProgram(self, text)

    def sync(self):
        pass

    def close(self):
        pass

    def reset_counters(self):
        self.launches = 0

    def counters(self):
        return {"kernel_launches": self.launches}

    def timer_begin(self):
        import time
        self._t0 = time.perf_counter()

    def timer_end(self):
        import time
        return 1e3 * (time.perf_counter() - self._t0)

    def profile_enable(self, on):
        pass

    def profile_read(self):
        return {"contract_small": {"ms": 0.5, "launches": 10, "bytes": 1e6, "flops": 1e6}}


def _run(monkeypatch, argv):
    lines = []
    monkeypatch.setattr(bench, "emit", lines.append)
    monkeypatch.setattr(sys, "argv", ["bench.py"] + argv)
    monkeypatch.setattr(os, "dup2", lambda a, b: None)   # keep pytest's capture intact
    bench.main()
    assert len(lines) == 1
    json.dumps(lines[0])   # serialisable
    return lines[0]


@pytest.mark.parametrize("workload,calls,macs", [("ghz3", 5, 92), ("qft10", 259, 729084)])
def test_config_reference_arm(monkeypatch, workload, calls, macs):
    line = _run(monkeypatch, ["--workload", workload, "--impl", "reference", "--steps", "1",
                              "--warmup", "0"])
    assert line["impl"] == "reference" and line["unit"] == "ms" and not line["higher_is_better"]
    assert line["config"]["contract_calls"] == calls      # 259 contractions (QFT-10)
    assert line["config"]["complex_macs"] == macs         # 7.3e5 MACs (QFT-10)
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["dtype"] == ("c64" if workload == "ghz3" else "c128")


@pytest.mark.parametrize("workload", ["ghz3", "qft10"])
def test_config_device_arm_control_flow(monkeypatch, workload):
    import torch
    import picoquant_jl_b200.host.b200_backend as bb
    monkeypatch.setattr(bb, "B200Backend", _FakeDevice)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    line = _run(monkeypatch, ["--workload", workload, "--steps", "2", "--warmup", "1"])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
                "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "e2e",
                "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert key in line, key
    assert line["gpu_launches"] == 2 * 7
    assert line["roofline"]["bound"] == "hbm" and line["roofline"]["frac"] > 0
    assert line["e2e"]["d2h_bytes_per_step"] == (8 * 8 if workload == "ghz3" else 1024 * 16)
    assert line["e2e"]["h2d_bytes_per_step"] > 0
    tol = 1e-5 if workload == "ghz3" else 1e-10
    assert line["cpu_baseline"]["parity_rel_l2_device_vs_oracle"] < tol


def test_stream_costs_matches_metrics_golden():
    """8 / 44 / 124 (test/layer2_tests.jl:127-132) are node-dims metrics; on this unsliced
    stream the data extents agree, so the MAC count of the stream equals Metrics.flops."""
    from picoquant_jl_b200.host import (Circuit, DSLBackend, add_input,
                                        convert_circuit_to_network,
                                        full_wavefunction_contraction)
    circ = Circuit(3)
    circ.h(0)
    circ.cx(0, 1)
    circ.cx(0, 1)
    circ.cx(0, 2)
    dsl = DSLBackend()
    tn = convert_circuit_to_network(circ, dsl)
    add_input(tn, "000")
    full_wavefunction_contraction(tn, "vector")
    costs = bench.stream_costs(dsl.text(), dsl.store, 16)
    assert costs["macs"] == dsl.metrics.flops == 124
    assert costs["contractions"] == 6 and costs["largest_elems"] == 16
