#!/usr/bin/env python
"""bench.py -- sliced random-quantum-circuit single-amplitude contraction on B200.

Workload (BASELINE.json config 5, the one the metric is quoted on): Google-style
RQC on a 7x7 grid, depth 24 (``create_RQC`` gate sequence, numpy seed 0), input
|0..0>, output bitstring 0..0, two-qubit gates SVD-split (``decompose=true``), network
sliced on the first k=6 virtual bonds in edge order (``slicing.jl:41-56``) into P=64
slices, each slice contracted with the harness's deterministic sweep plan through
``contract_network!``'s command stream.  A "step" is one full amplitude: all P slices
(sharded contiguously over the N GPUs), device-side accumulation, one NCCL all-reduce.

    python bench.py --gpus N --steps K --warmup W            # this framework
    python bench.py --impl reference --gpus N --steps K ...  # CPU oracle arm

Prints ONE JSON line (rank 0).  ``value`` = whole-job real TFLOP/s of contraction
(8 flops per complex MAC, MACs counted on data extents) with inputs resident in HBM;
``e2e`` = the same metric through the backend API from host buffers (H2D of every
gate tensor + compile-free replay + D2H of the amplitude inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import picoquant_jl_b200  # noqa: E402,F401
from picoquant_jl_b200.host import create_RQC  # noqa: E402
from picoquant_jl_b200.host.backends import TensorStore, parse_dsl  # noqa: E402
from picoquant_jl_b200.host.planner import sweep_plan  # noqa: E402
from picoquant_jl_b200.host.sliced import (SlicedContraction, partitions_of_rank,  # noqa: E402
                                           record_sliced_contraction)

METRIC = "rqc_amplitude_contraction_tflops"
UNIT = "TFLOP/s"
GEMM_CLASSES = ("gemm_tensor", "gemm_simt", "gemm_int8")

# BASELINE.json configs: name -> (config number, description)
WORKLOADS = {
    "ghz3": (1, "GHZ-3 full wave function from ghz_3.qasm (bin/contract_qasm.jl flow)"),
    "qft10": (2, "QFT-10 full wave function from qft_10.qasm"),
    "qft26": (3, "QFT-26 full wave function, create_qft_circuit(26), 2^26 amplitudes"),
    "rqc6x6": (4, "RQC 6x6 depth 20 single amplitude, un-decomposed network, greedy plan"),
    "rqc7x7_sliced": (5, "sliced RQC 7x7 depth 24 single amplitude (dist_slicing_example.jl path)"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default=None, choices=["c128", "c64"],
                    help="backend element type (default: c64 for ghz3 like bin/contract_qasm.jl, "
                         "c128 otherwise)")
    ap.add_argument("--workload", default="rqc7x7_sliced", choices=sorted(WORKLOADS),
                    help="rqc7x7_sliced = BASELINE config 5 (the headline, default); the others "
                         "are BASELINE configs 1-4, single GPU, same JSON contract")
    ap.add_argument("--rows", type=int, default=7)
    ap.add_argument("--cols", type=int, default=7)
    ap.add_argument("--depth", type=int, default=24)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--slices", type=int, default=64)
    ap.add_argument("--lanes", type=int, default=4,
                    help="slices kept in flight per GPU (pq_program_run_slices)")
    ap.add_argument("--cpu-sample-slices", type=int, default=2)
    ap.add_argument("--ozaki", type=int, default=0, choices=[0, 4, 6],
                    help="force every eligible GEMM step (K, N <= 64) onto the INT8 tensor-core kernel: 6 with "
                         "--dtype c128 (option zgemm_ozaki), 4 with --dtype c64 (option cgemm_ozaki); default: "
                         "the library's own policy (ozaki_auto)")
    ap.add_argument("--no-int8", action="store_true",
                    help="option ozaki_auto = 0: keep the skinny GEMM steps on DMMA / tcgen05 3xTF32 (A/B)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    a = ap.parse_args()
    if a.dtype is None:
        a.dtype = "c64" if a.workload == "ghz3" else "c128"
    return a


def build_workload(a):
    circ = create_RQC(a.rows, a.cols, a.depth, seed=a.seed)
    n = circ.n_qubits

    def plan_fn(tn, sliced):
        return sweep_plan(tn, a.rows, a.cols, sliced_bonds=sliced)

    rec = record_sliced_contraction(circ, a.slices, 1, plan_fn=plan_fn, output_config="0" * n)
    name = "rqc_%dx%d_d%d_seed%d_amplitude_sliced_P%d" % (a.rows, a.cols, a.depth, a.seed, a.slices)
    return circ, rec, name


def workload_config(a, name, rec, macs):
    """The `config` object of the JSON line: the WORKLOAD only, identical for the GPU arm and the
    reference (CPU) arm so that the driver can compare them; how an arm executes it goes into
    its own `run` object."""
    itemsize = 16 if a.dtype == "c128" else 8
    return {"workload": name, "baseline_config": 5, "slices": a.slices, "plan": "sweep (harness planner)",
            "complex_macs_per_slice": macs,
            "contract_calls_per_slice": sum(1 for c, _ in parse_dsl(rec.text) if c == "ncon"),
            "largest_intermediate_elems": 1 << 24,
            "l2": "inputs larger than L2: per-step intermediates of 2^24 elements (%d MiB) stream "
                  "through the 126 MB L2" % ((itemsize << 24) >> 20)}


def slice_macs(rec):
    """Complex MACs of one slice on DATA extents (sliced bonds have extent 1)."""
    shapes, macs = {}, 0
    for cmd, x in parse_dsl(rec.text):
        if cmd == "tensor":
            shapes[x["t"]] = list(rec.store.read(x["key"]).shape)
        elif cmd == "view":
            s = list(shapes[x["t"]])
            s[x["axis"] - 1] = len(x["idx"])
            shapes[x["v"]] = s
        elif cmd == "ncon":
            sa, sb = shapes[x["A"]], shapes[x["B"]]
            bset = set(x["b_idx"])
            k = 1
            for d, lab in zip(sa, x["a_idx"]):
                if lab in bset:
                    k *= d
            m = int(np.prod(sa)) // k
            nn = int(np.prod(sb)) // k
            macs += m * nn * k
            aset = set(x["a_idx"])
            shapes[x["C"]] = [d for d, lab in zip(sa, x["a_idx"]) if lab not in bset] + \
                             [d for d, lab in zip(sb, x["b_idx"]) if lab not in aset]
        elif cmd == "del":
            shapes.pop(x["t"], None)
    return macs


# ---------------------------------------------------------------------------
# CPU arm: the oracle (NumPy/OpenBLAS restatement of the reference's CPU path)
# ---------------------------------------------------------------------------
def run_cpu_slices(rec, dtype, partitions):
    from oracle.interactive import execute_dsl
    total = 0
    for p in partitions:
        out = TensorStore()
        execute_dsl(rec.text_for(p), rec.store, dtype, output_store=out)
        total = total + out.read("result")
    return total


def use_all_host_threads():
    """The CPU arm uses every host core (torchrun exports OMP_NUM_THREADS=1, which would
    silently make OpenBLAS single-threaded).  Returns the BLAS thread count in effect."""
    n = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=n)
        return max([i.get("num_threads", 1) for i in threadpool_info()] + [1])
    except Exception:  # noqa: BLE001
        return n


def host_cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return "%s x %d" % (line.split(":", 1)[1].strip(), os.cpu_count() or 1)
    except OSError:
        pass
    return "unknown x %d" % (os.cpu_count() or 1)


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([i.get("num_threads", 1) for i in threadpool_info()] + [1])
    except Exception:  # noqa: BLE001
        return os.cpu_count() or 1


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    use_all_host_threads()
    circ, rec, name = build_workload(a)
    dtype = np.complex128 if a.dtype == "c128" else np.complex64
    macs = slice_macs(rec)
    per_step = 1  # one slice per step keeps K+W steps within minutes
    for w in range(a.warmup):
        run_cpu_slices(rec, dtype, [1 + (w % a.slices)])
    t0 = time.perf_counter()
    for s in range(a.steps):
        run_cpu_slices(rec, dtype, [1 + (s % a.slices)])
    dt = (time.perf_counter() - t0) / max(1, a.steps)
    value = 8.0 * macs * per_step / dt / 1e12
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
        "config": workload_config(a, name, rec, macs),
        "run": {"note": "reference CPU path restated in NumPy/OpenBLAS (the Julia reference cannot run "
                        "here); each step contracts ONE of the %d slices -- a bounded sample of the "
                        "workload, the metric is a rate; a full amplitude costs %d x this"
                        % (a.slices, a.slices),
                "slices_per_step": per_step, "host_cpu": host_cpu_model()},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
                         "sample": "1 slice of %d per step" % a.slices},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "amplitude_wall_ms_extrapolated": dt * 1e3 * a.slices,
    }
    emit(line)


# ---------------------------------------------------------------------------
# BASELINE configs 1-4 (single GPU): --workload ghz3 | qft10 | qft26 | rqc6x6
# ---------------------------------------------------------------------------
def stream_costs(text, store, itemsize):
    """Algorithmic units of a command stream on data extents: complex MACs,
    bytes = (MK + KN + MN) * sizeof per ``ncon`` + 2 * numel * sizeof per non-identity
    ``permute`` and per ``view``, the number of contractions and the largest tensor."""
    shapes, macs, nbytes, ncon, biggest = {}, 0, 0, 0, 0
    for cmd, x in parse_dsl(text):
        if cmd == "tensor":
            shapes[x["t"]] = list(store.read(x["key"]).shape)
        elif cmd == "view":
            sh = list(shapes[x["t"]])
            sh[x["axis"] - 1] = len(x["idx"])
            shapes[x["v"]] = sh
            nbytes += 2 * int(np.prod(sh)) * itemsize
        elif cmd == "permute":
            sh = shapes[x["t"]]
            if list(x["axes"]) != list(range(1, len(sh) + 1)):
                nbytes += 2 * int(np.prod(sh)) * itemsize
            shapes[x["t"]] = [sh[i - 1] for i in x["axes"]]
        elif cmd == "reshape":
            sh = shapes[x["t"]]
            shapes[x["t"]] = [int(np.prod([sh[i - 1] for i in g])) for g in x["groups"]]
        elif cmd == "ncon":
            sa, sb = shapes[x["A"]], shapes[x["B"]]
            aset, bset = set(x["a_idx"]), set(x["b_idx"])
            k = 1
            for d, lab in zip(sa, x["a_idx"]):
                if lab in bset:
                    k *= d
            m = int(np.prod(sa)) // k
            nn = int(np.prod(sb)) // k
            macs += m * nn * k
            nbytes += (m * k + k * nn + m * nn) * itemsize
            ncon += 1
            biggest = max(biggest, m * nn, m * k, nn * k)
            shapes[x["C"]] = [d for d, lab in zip(sa, x["a_idx"]) if lab not in bset] + \
                             [d for d, lab in zip(sb, x["b_idx"]) if lab not in aset]
        elif cmd == "del":
            shapes.pop(x["t"], None)
    return {"macs": macs, "bytes": nbytes, "contractions": ncon, "largest_elems": biggest}


class ConfigWorkload:
    """One of BASELINE configs 1-4: the circuit, the reference-facing call sequence a user
    makes (``flow``) and that sequence recorded as a ``.tl`` command stream."""

    def __init__(self, workload, seed=0, qft_n=26):
        from picoquant_jl_b200.host import (DSLBackend, add_input, add_output, contract_network,
                                            convert_circuit_to_network, create_qft_circuit,
                                            full_wavefunction_contraction, load_qasm_as_circuit)
        from picoquant_jl_b200.host.planner import greedy_plan
        golden = os.path.join(ROOT, "tests", "golden")
        self.amplitude = workload == "rqc6x6"
        if workload == "ghz3":
            with open(os.path.join(golden, "ghz_3.qasm")) as f:
                self.circ = load_qasm_as_circuit(f.read())
            self.name = "ghz_3.qasm_full_wavefunction"
        elif workload == "qft10":
            with open(os.path.join(golden, "qft_10.qasm")) as f:
                self.circ = load_qasm_as_circuit(f.read())
            self.name = "qft_10.qasm_full_wavefunction"
        elif workload == "qft26":
            self.circ = create_qft_circuit(qft_n)
            self.name = "qft_%d_full_wavefunction" % qft_n
        elif workload == "rqc6x6":
            self.circ = create_RQC(6, 6, 20, seed=seed)
            self.name = "rqc_6x6_d20_seed%d_amplitude" % seed
        else:
            raise ValueError(workload)
        n = self.n = self.circ.n_qubits
        self.plan = None

        def flow(backend):
            tn = convert_circuit_to_network(self.circ, backend)
            add_input(tn, "0" * n)
            if self.amplitude:
                add_output(tn, "0" * n)
                if self.plan is None:
                    self.plan = greedy_plan(tn)
                contract_network(tn, self.plan, "")
            else:
                full_wavefunction_contraction(tn, "vector")
            return tn

        self.flow = flow
        dsl = DSLBackend()
        flow(dsl)
        self.text, self.store = dsl.text(), dsl.store

    def h2d_bytes(self, itemsize):
        return sum(int(self.store.read(a["key"]).size) * itemsize
                   for cmd, a in parse_dsl(self.text) if cmd == "tensor")


def cpu_flow_seconds(w, dtype, min_seconds=1.0, max_reps=200):
    """The reference-facing flow through the oracle backend, repeated until ``min_seconds``
    of CPU time have been spent; returns (seconds per flow, repetitions, result)."""
    from oracle.interactive import OracleBackend
    use_all_host_threads()
    w.flow(OracleBackend(dtype))   # warm-up
    reps, t0 = 0, time.perf_counter()
    while True:
        ob = OracleBackend(dtype)
        w.flow(ob)
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds or reps >= max_reps:
            break
    return dt / reps, reps, np.asarray(ob.load_tensor_data("result"))


def cpu_config_sample(a, w, dtype):
    """CPU figure for a config: the whole flow for configs 1, 2 and 4; for QFT-26 (minutes
    of NumPy transposes of a 1 GiB state per pass) QFT-22 timed and scaled by the ratio of
    algorithmic bytes -- every step of either circuit is a bandwidth-bound gate application."""
    itemsize = np.dtype(dtype).itemsize
    if a.workload == "qft26":
        small = ConfigWorkload("qft26", qft_n=22)
        sec, reps, _ = cpu_flow_seconds(small, dtype, min_seconds=0.0, max_reps=1)
        scale = stream_costs(w.text, w.store, itemsize)["bytes"] / \
            stream_costs(small.text, small.store, itemsize)["bytes"]
        return sec * scale, "QFT-22 flow (%.1f s) x %.1f (ratio of algorithmic bytes)" % (sec, scale), None
    sec, reps, res = cpu_flow_seconds(w, dtype)
    return sec, "whole flow, mean of %d repetitions" % reps, res


def config_line(a, w, costs, ms, impl=None):
    """Metric fields of a config line: wall time for the full-wave-function configs, the
    headline TFLOP/s metric for the single-amplitude RQC."""
    if w.amplitude:
        return {"metric": METRIC, "value": 8.0 * costs["macs"] / (ms * 1e-3) / 1e12, "unit": UNIT,
                "higher_is_better": True}
    return {"metric": "full_wavefunction_contraction_wall_ms", "value": ms, "unit": "ms",
            "higher_is_better": False}


def config_arm(a):
    """BASELINE configs 1-4 with the same JSON contract as the headline workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return   # single-GPU configurations ("replicas only")
    dtype = np.complex128 if a.dtype == "c128" else np.complex64
    itemsize = np.dtype(dtype).itemsize
    w = ConfigWorkload(a.workload, seed=a.seed)
    costs = stream_costs(w.text, w.store, itemsize)
    cfg = {"workload": w.name, "baseline_config": WORKLOADS[a.workload][0],
           "contract_calls": costs["contractions"], "complex_macs": costs["macs"],
           "algorithmic_bytes": costs["bytes"], "largest_tensor_elems": costs["largest_elems"]}
    base = {"n_gpus": 1, "steps": a.steps, "warmup": a.warmup, "scaling": "strong",
            "vs_baseline": None, "dtype": a.dtype, "data": "synthetic"}

    if a.impl == "reference":
        for _ in range(a.warmup if a.workload != "qft26" else 0):
            cpu_config_sample(a, w, dtype)
        secs = []
        for _ in range(a.steps if a.workload != "qft26" else 1):
            sec, sample, _ = cpu_config_sample(a, w, dtype)
            secs.append(sec)
        ms = 1e3 * float(np.mean(secs))
        line = dict(base)
        line.update(config_line(a, w, costs, ms))
        line.update({"impl": "reference", "ms_per_step": ms, "config": cfg,
                     "cpu_baseline": {"value": line["value"], "unit": line["unit"],
                                      "cores": cpu_threads(), "kind": "port", "sample": sample},
                     "e2e": {"value": line["value"], "unit": line["unit"],
                             "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        emit(line)
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    from picoquant_jl_b200.host.b200_backend import B200Backend
    b = B200Backend(dtype, device=0)
    if a.no_int8:
        b.set_option("ozaki_auto", 0)
        cfg["ozaki_auto"] = 0
    if a.ozaki:   # forced INT8 tensor-core GEMMs (incl. the experimental K-looped kernel)
        if (a.dtype == "c128") != (a.ozaki == 6):
            raise SystemExit("--ozaki 6 goes with --dtype c128, --ozaki 4 with --dtype c64")
        b.set_option("zgemm_ozaki" if a.dtype == "c128" else "cgemm_ozaki", a.ozaki)
        cfg["ozaki_groups"] = a.ozaki
    for cmd, x in parse_dsl(w.text):
        if cmd == "tensor":
            b.save_tensor_data(x["key"], w.store.read(x["key"]))
    prog = b.compile_program(w.text)
    assert prog.macs == costs["macs"], (prog.macs, costs["macs"])
    for _ in range(max(3, a.warmup)):
        prog.run()
    b.sync()
    sampler = ClockSampler(0)
    sampler.start()
    b.reset_counters()
    b.timer_begin()
    for _ in range(a.steps):
        prog.run()
    ms = b.timer_end() / a.steps
    launches = b.counters()["kernel_launches"]
    clocks = sampler.stop()
    result = np.asarray(b.load_tensor_data("result")).reshape(-1)

    # end to end: the reference-facing calls from host arrays on a fresh backend (upload of
    # every gate tensor, one backend call per contraction, D2H of the result)
    e2e_times, d2h = [], 0
    for it in range(3):
        be = B200Backend(dtype, device=0)
        be.sync()
        t0 = time.perf_counter()
        w.flow(be)
        res_e2e = np.asarray(be.load_tensor_data("result")).reshape(-1)
        e2e_times.append(time.perf_counter() - t0)
        d2h = int(res_e2e.size) * itemsize
        be.close()
    e2e_ms = 1e3 * float(np.mean(e2e_times[1:]))

    # per-kernel split of one eager, event-timed pass
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            hbm, hbm_src = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json (driver-written)"
    except Exception:  # noqa: BLE001
        hbm, hbm_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    b.profile_enable(True)
    prog.run()
    prof = b.profile_read()
    b.profile_enable(False)
    kernels = {c: {"launches": v["launches"], "ms": v["ms"],
                   "avg_launch_us": 1e3 * v["ms"] / v["launches"],
                   "achieved_gbs": v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["bytes"] else None,
                   "achieved_tflops": v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["flops"] else None}
               for c, v in prof.items()}
    dom = max(prof, key=lambda c: prof[c]["ms"])
    if w.amplitude and dom in ("gemm_tensor", "gemm_simt"):
        tdt = torch.complex128 if a.dtype == "c128" else torch.complex64
        x = torch.randn(4096, 4096, dtype=tdt, device="cuda")
        torch.matmul(x, x)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e30
        for _ in range(3):
            e0.record()
            torch.matmul(x, x)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        lib = 8.0 * 4096 ** 3 / (best * 1e-3) / 1e12
        ach = kernels[dom]["achieved_tflops"]
        if a.dtype == "c128":
            peak, src = lib, "cuBLAS ZGEMM 4096^3 measured in this run (the FP64 DMMA pipe)"
        else:
            peak = tf32_split_peak_tflops()
            src = ("TF32 tensor pipe / 3: cuBLAS TF32 SGEMM 8192^3 measured in this run, divided by the 3 pipe "
                   "flops a 3xTF32 split product spends per algorithmic flop")
        roofline = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": peak,
                    "unit": "TFLOP/s", "frac": ach / peak, "traffic": None, "peak_source": src,
                    "library_gemm_tflops": lib,
                    "library_gemm": "cuBLAS %s 4096^3" % ("ZGEMM" if a.dtype == "c128" else "CGEMM")}
    else:
        # whole-stream figure: algorithmic bytes of every step / device time of the replay
        ach = costs["bytes"] / (ms * 1e-3) / 1e9
        roofline = {"kernel": "whole stream (dominant class: %s)" % dom, "bound": "hbm",
                    "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                    "traffic": None, "peak_source": hbm_src}

    cpu = None
    if not a.no_cpu_baseline:
        sec, sample, ref = cpu_config_sample(a, w, dtype)
        cl = config_line(a, w, costs, 1e3 * sec)
        cpu = {"value": cl["value"], "unit": cl["unit"], "cores": cpu_threads(), "kind": "port",
               "sample": sample + ", NumPy/OpenBLAS oracle"}
        if ref is not None:
            ref = ref.reshape(-1)
            err = float(np.linalg.norm(result - ref) / np.linalg.norm(ref))
        else:   # QFT|0..0> = uniform superposition (closed form)
            err = float(np.linalg.norm(result - 2.0 ** (-w.n / 2)) / 1.0)
        cpu["parity_rel_l2_device_vs_oracle"] = err
        tol = 1e-10 if a.dtype == "c128" else 1e-5
        if not err < tol:
            raise SystemExit("parity failure on %s: rel-L2 %g" % (w.name, err))

    working_set = costs["largest_elems"] * itemsize
    cfg.update({"kernel_launches_per_step": prog.launches, "arena_bytes": prog.arena_bytes,
                "us_per_contract_call": 1e3 * ms / max(1, costs["contractions"]),
                "l2": ("tensors larger than L2 (largest %d MiB)" % (working_set >> 20))
                if working_set > (126 << 20) else
                "working set (largest tensor %d KiB) fits L2: this configuration is launch / "
                "latency bound by construction; no flush between steps" % (working_set >> 10)})
    line = dict(base)
    line.update(config_line(a, w, costs, ms))
    e2e_line = config_line(a, w, costs, e2e_ms)
    line.update({"ms_per_step": ms, "config": cfg,
                 "e2e": {"value": e2e_line["value"], "unit": e2e_line["unit"],
                         "h2d_bytes_per_step": w.h2d_bytes(itemsize), "d2h_bytes_per_step": d2h,
                         "ms_per_step": e2e_ms,
                         "note": "eager backend calls from host arrays on a fresh handle, "
                                 "Python host mirror included"},
                 "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
                 "kernels": kernels, "cpu_baseline": cpu})
    emit(line)
    prog.close()
    b.close()


# ---------------------------------------------------------------------------
# clocks sampler (nvidia-smi, during the timed region)
# ---------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm),
                "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
_REAL_STDOUT = None



def tf32_split_peak_tflops():
    """Roofline denominator of the ComplexF32 tcgen05 GEMM: the TF32 tensor-pipe rate
    measured on this box (cuBLAS SGEMM with TF32 inputs, 8192^3) divided by 3 -- a complex
    product costs 12 TF32 MMAs (hi*hi + hi*lo + lo*hi for each of its four real products) against
    4 on exact inputs, so 3 pipe flops are spent per algorithmic flop."""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        x = torch.randn(n, n, dtype=torch.float32, device="cuda")
        y = torch.randn(n, n, dtype=torch.float32, device="cuda")
        torch.matmul(x, y)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e30
        for _ in range(3):
            e0.record()
            torch.matmul(x, y)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12 / 3.0
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old

def emit(line: dict) -> None:
    """Prints the ONE JSON line on the real stdout (native libraries such as NCCL write
    their banners to fd 1, which is pointed at stderr for the duration of the run)."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    a = parse_args()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if a.workload != "rqc7x7_sliced":
        config_arm(a)
        return
    if a.impl == "reference":
        reference_arm(a)
        return

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from picoquant_jl_b200.host.b200_backend import B200Backend
    dtype = np.complex128 if a.dtype == "c128" else np.complex64
    circ, rec, name = build_workload(a)
    macs = slice_macs(rec)
    mine = partitions_of_rank(a.slices, rank, world)

    b = B200Backend(dtype, device=local_rank)
    if a.no_int8:
        b.set_option("ozaki_auto", 0)
    if a.ozaki:
        if (a.dtype == "c128") != (a.ozaki == 6):
            raise SystemExit("--ozaki 6 goes with --dtype c128, --ozaki 4 with --dtype c64")
        b.set_option("zgemm_ozaki" if a.dtype == "c128" else "cgemm_ozaki", a.ozaki)
    if world > 1:
        ids = [B200Backend.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        b.comm_init(ids[0], rank, world)
    sc = SlicedContraction(b, rec)          # uploads the gate tensors, compiles the plan
    assert sc.program.macs == macs, (sc.program.macs, macs)

    def barrier():
        b.sync()
        if world > 1:
            dist.barrier()

    def amplitude_step():
        b.delete_tensor("partial_sum")
        sc.run(mine, "partial_sum", lanes=a.lanes)
        if world > 1:
            b.allreduce_sum("partial_sum")

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(a.warmup):
        amplitude_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    b.reset_counters()
    b.timer_begin()
    for _ in range(a.steps):
        amplitude_step()
    ms = b.timer_end()
    barrier()
    launches = b.counters()["kernel_launches"]
    clocks = sampler.stop() if rank == 0 else None
    ms = max_over_ranks(ms)
    ms_per_step = ms / a.steps
    total_flops = 8.0 * macs * a.slices
    value = total_flops / (ms_per_step * 1e-3) / 1e12
    amplitude = b.load_tensor_data("partial_sum")

    # ---- the same amplitude with slice-invariant hoisting (reported separately) -----
    # pq_program_prepare executes the steps that do not depend on the slice once per
    # amplitude; only the slice-dependent steps are replayed per slice.  Not the headline:
    # `value` above executes the full stream for every slice, like the reference flow.
    def amplitude_step_hoisted():
        b.delete_tensor("partial_sum")
        sc.run(mine, "partial_sum", hoist=True, lanes=a.lanes)
        if world > 1:
            b.allreduce_sum("partial_sum")

    for _ in range(a.warmup):
        amplitude_step_hoisted()
    barrier()
    b.timer_begin()
    for _ in range(a.steps):
        amplitude_step_hoisted()
    ms_h = max_over_ranks(b.timer_end()) / a.steps
    barrier()
    amp_h = b.load_tensor_data("partial_sum")
    executed_h = 8.0 * world * (sc.program.macs_invariant + len(mine) * sc.program.macs_dependent)
    hoisted = {
        "ms_per_step": ms_h,
        "effective_tflops": total_flops / (ms_h * 1e-3) / 1e12,
        "executed_tflops": executed_h / (ms_h * 1e-3) / 1e12,
        "invariant_mac_share": sc.program.macs_invariant / max(1, sc.program.macs),
        "amplitude_rel_diff_vs_full_replay": float(abs(amp_h - amplitude) / abs(amplitude)),
        "note": "slice-invariant steps executed once per amplitude (per GPU), slice-dependent "
                "steps per slice; effective = reference-equivalent flops / time",
    }
    sc.program.set_hoist(False)

    # ---- end to end through the backend API, from host buffers ------------------
    h2d = 0
    e2e_times = []
    for it in range(max(2, min(a.steps, 3)) + 1):
        barrier()
        t0 = time.perf_counter()
        h2d = sc.upload()                      # H2D of every gate tensor (host arrays)
        amplitude_step()
        amp_e2e = b.load_tensor_data("partial_sum")   # D2H of the amplitude (syncs)
        dt = time.perf_counter() - t0
        if it > 0:
            e2e_times.append(max_over_ranks(dt))
    e2e_s = float(np.mean(e2e_times))
    e2e_value = total_flops / e2e_s / 1e12
    d2h = int(np.asarray(amp_e2e).size * np.dtype(dtype).itemsize)

    # ---- per-kernel roofline, measured inside the graph replays of the timed path -----
    kernels, roofline = {}, None
    oz_groups = getattr(a, "ozaki", 0)
    peaks = {}
    if rank == 0 and not a.no_profile:
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                mp = json.load(f)
            peaks["hbm_gbs"] = float(mp["hbm_gbs"])
            peaks["hbm_source"] = "MEASURED_PEAKS.json (driver-written, of measured)"
        except Exception:  # noqa: BLE001
            peaks["hbm_gbs"] = 6650.0
            peaks["hbm_source"] = "fallback 6.65 TB/s (B200_PROFILING.md)"
        peaks["fp64_dmma_probe_tflops"] = b.microbench("dmma_tflops")
        try:
            n = 4096
            tdt = torch.complex128 if a.dtype == "c128" else torch.complex64
            x = torch.randn(n, n, dtype=tdt, device="cuda")
            y = torch.randn(n, n, dtype=tdt, device="cuda")
            torch.matmul(x, y)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = 1e30
            for _ in range(3):
                e0.record()
                torch.matmul(x, y)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            peaks["cublas_zgemm_tflops"] = 8.0 * n ** 3 / (best * 1e-3) / 1e12
            del x, y
            x = torch.randn(1 << 18, 64, dtype=tdt, device="cuda")
            y = torch.randn(64, 64, dtype=tdt, device="cuda")
            torch.matmul(x, y)
            best = 1e30
            for _ in range(3):
                e0.record()
                torch.matmul(x, y)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            peaks["cublas_zgemm_skinny_tflops"] = 8.0 * (1 << 18) * 64 * 64 / (best * 1e-3) / 1e12
            peaks["cublas_zgemm_skinny_shape"] = "M=2^18 N=64 K=64, operands already in GEMM layout"
            del x, y
        except Exception as e:  # noqa: BLE001
            peaks["cublas_zgemm_tflops"] = None
            peaks["cublas_error"] = repr(e)
        # ---- in-graph profile: the SAME replays as the timed loop (same lanes, same parallel
        # branches), each kernel node bracketed by event-record nodes.  busy = union of a
        # class's kernel intervals, so concurrent kernels are not counted twice and
        # busy <= wall by construction.
        n_prof = min(len(mine), 3 * max(1, a.lanes))
        starts = [rec.view_starts(p) for p in mine[:n_prof]] if rec.bond_labels else [[]] * n_prof
        gp = sc.program.profile_slices(starts, lanes=a.lanes)
        wall_prof = gp["wall_ms"]
        prof = gp["classes"]
        for cls, v in prof.items():
            k = {"launches_per_slice": v["launches"] / n_prof,
                 "busy_ms_per_slice": v["busy_ms"] / n_prof,
                 "sum_ms_per_slice": v["sum_ms"] / n_prof,
                 "avg_launch_us": 1e3 * v["sum_ms"] / v["launches"],
                 "share_of_profiled_wall": v["busy_ms"] / wall_prof if wall_prof else None}
            if v["flops"] and cls in GEMM_CLASSES:
                k["achieved_tflops"] = v["flops"] / (v["busy_ms"] * 1e-3) / 1e12
            if v["bytes"]:
                k["achieved_gbs"] = v["bytes"] / (v["busy_ms"] * 1e-3) / 1e9
            kernels[cls] = k
        timed_slice_ms = ms_per_step / max(1, len(mine))
        kernels["_profile"] = {
            "how": "pq_program_profile_slices: event-record nodes inside the graph replays, "
                   "%d slices on %d lanes; busy = union of the class's kernel intervals" % (n_prof, a.lanes),
            "profiled_wall_ms_per_slice": wall_prof / n_prof,
            "timed_ms_per_slice": timed_slice_ms,
            "note": "the profiled window holds the pipeline fill of its first slices and the "
                    "event nodes; the timed loop does not"}
        dom = max(prof, key=lambda c: prof[c]["busy_ms"])
        for cls in kernels:
            if cls in prof:
                # what the class holds of ONE TIMED slice: its busy time per slice over the timed
                # slice time (must be <= 1 up to the profile's own overhead)
                kernels[cls]["share_of_timed_slice"] = (prof[cls]["busy_ms"] / n_prof) / timed_slice_ms
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
            ent = tj.get("%s/%s" % (dom, a.dtype))
            if ent:
                traffic, traffic_src = ent["dram_bytes_per_launch"], ent["source"]
        except Exception:  # noqa: BLE001
            pass
        share = kernels[dom].get("share_of_timed_slice")
        if dom == "gemm_int8":
            # INT8 Ozaki GEMM steps: judged against whichever roof they are closer to -- HBM
            # (algorithmic operand + result bytes) or the INT8 tensor pipe, whose complex-flop
            # equivalent is the measured kind::i8 issue rate over the int8 MACs one complex MAC
            # costs (4 real products x digit-plane pairs)
            pairs = 21 if a.dtype == "c128" else 10   # digit-plane pairs: 6 / 4 accumulator groups
            tops = b.microbench("umma_i8_tops_n64")
            peaks["umma_i8_tops_n64"] = tops
            tensor_peak = tops * 8.0 / (2.0 * 4.0 * pairs)
            f_t = kernels[dom]["achieved_tflops"] / tensor_peak
            f_h = kernels[dom]["achieved_gbs"] / peaks["hbm_gbs"]
            common = {"kernel": "gemm_int8 (k_ozaki_t, tcgen05.mma kind::i8)", "traffic": traffic, "traffic_source": traffic_src,
                      "share_of_timed_slice": share, "frac_hbm": f_h, "frac_int8_pipe": f_t,
                      "achieved_tflops_algorithmic": kernels[dom]["achieved_tflops"]}
            if f_h >= f_t:
                roofline = dict(common, bound="hbm", achieved=kernels[dom]["achieved_gbs"],
                                peak=peaks["hbm_gbs"], unit="GB/s", frac=f_h,
                                peak_source=peaks["hbm_source"],
                                note="algorithmic (MK + KN + MN) * sizeof(T) bytes of the class / its "
                                     "in-graph busy time")
            else:
                roofline = dict(common, bound="tensor", achieved=kernels[dom]["achieved_tflops"],
                                peak=tensor_peak, unit="TFLOP/s", frac=f_t,
                                peak_source="measured kind::i8 issue rate %.0f TOPS / %d int8 MACs per "
                                            "complex MAC" % (tops, 4 * pairs))
        elif dom in GEMM_CLASSES:
            peak = peaks.get("cublas_zgemm_tflops") or peaks["fp64_dmma_probe_tflops"]
            if a.dtype == "c64":
                peaks["tf32_pipe_over_3_tflops"] = peak = tf32_split_peak_tflops()
            ach = kernels[dom]["achieved_tflops"]
            roofline = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": peak,
                        "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
                        "traffic_source": traffic_src, "share_of_timed_slice": share,
                        "note": ("achieved = algorithmic 8*M*N*K flops of the class / its in-graph "
                                 "busy time; the skinny kernels form a complex product from three "
                                 "DMMAs (3M), i.e. issue 6*M*N*K pipe flops: frac_of_3m_pipe is the "
                                 "same figure against the DMMA issue probe x 8/6" if a.dtype == "c128" else
                                 "achieved = algorithmic 8*M*N*K flops / in-graph busy time "
                                 "(tcgen05 3xTF32: 12 TF32 MMAs per complex product)"),
                        "peak_source": ("cuBLAS ZGEMM 4096^3 measured in this run "
                                        "(MEASURED_PEAKS.json has no FP64 / complex figure); the "
                                        "same library reaches %s TFLOP/s on the dominant skinny "
                                        "shape" % peaks.get("cublas_zgemm_skinny_tflops")) if a.dtype == "c128" else
                                       ("TF32 tensor pipe / 3 (cuBLAS TF32 SGEMM 8192^3 measured in this run; "
                                        "cuBLAS CGEMM 4096^3 reaches %s TFLOP/s)" % peaks.get("cublas_zgemm_tflops"))}
            if a.dtype == "c128":
                roofline["frac_of_3m_pipe"] = ach / (peaks["fp64_dmma_probe_tflops"] * 8.0 / 6.0)
        else:
            ach = kernels[dom]["achieved_gbs"]
            roofline = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"],
                        "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "traffic": traffic,
                        "traffic_source": traffic_src, "share_of_timed_slice": share,
                        "peak_source": peaks["hbm_source"]}
        # every GEMM-shaped class also against HBM (the K <= 32 / N = 8 sweep steps are HBM-bound)
        for cls in GEMM_CLASSES:
            if cls in kernels and "achieved_gbs" in kernels[cls]:
                kernels[cls]["frac_of_hbm"] = kernels[cls]["achieved_gbs"] / peaks["hbm_gbs"]
        if "permute_tiled" in kernels:
            kernels["permute_tiled"]["frac_of_hbm"] = kernels["permute_tiled"]["achieved_gbs"] / peaks["hbm_gbs"]
        if "gemm_tensor" in kernels and peaks.get("cublas_zgemm_tflops"):
            kernels["gemm_tensor"]["frac_of_cublas_zgemm"] = (kernels["gemm_tensor"]["achieved_tflops"]
                                                              / peaks["cublas_zgemm_tflops"])

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only) -----------------
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        sample = list(range(1, a.cpu_sample_slices + 1))
        use_all_host_threads()
        run_cpu_slices(rec, dtype, [1])  # warm-up (BLAS threads, page faults)
        t0 = time.perf_counter()
        part = run_cpu_slices(rec, dtype, sample)
        dt = time.perf_counter() - t0
        cpu = {"value": 8.0 * macs * len(sample) / dt / 1e12, "unit": UNIT, "cores": cpu_threads(),
               "kind": "port", "sample": "%d of %d slices (%.1f s), NumPy/OpenBLAS oracle"
                                         % (len(sample), a.slices, dt),
               "amplitude_wall_s_extrapolated": dt / len(sample) * a.slices}
        # one slice on a single BLAS thread, for context; never fatal
        try:
            from threadpoolctl import threadpool_limits
            with threadpool_limits(limits=1):
                t0 = time.perf_counter()
                run_cpu_slices(rec, dtype, [1])
                dt1 = time.perf_counter() - t0
            cpu["single_thread"] = {"value": 8.0 * macs / dt1 / 1e12, "unit": UNIT, "cores": 1,
                                    "sample": "1 of %d slices (%.1f s)" % (a.slices, dt1)}
            use_all_host_threads()
        except Exception as e:  # noqa: BLE001
            cpu["single_thread"] = {"error": repr(e)}
        # cheap parity guard: device partial sum of the same slices vs the ComplexF64 oracle
        # (through the same lanes configuration as the timed loop)
        b.delete_tensor("check_sum")
        sc.run(sample, "check_sum", lanes=a.lanes)
        dev = b.load_tensor_data("check_sum")
        ref64 = part if a.dtype == "c128" else run_cpu_slices(rec, np.complex128, sample)
        err = abs(dev - ref64) / abs(ref64)
        cpu["parity_rel_err_device_vs_f64_oracle"] = float(err)
        if a.dtype == "c64":   # what the CPU path itself loses in ComplexF32 on this scalar
            cpu["rel_err_c64_oracle_vs_f64_oracle"] = float(abs(part - ref64) / abs(ref64))
        tol = 1e-10 if a.dtype == "c128" else 1e-5
        if not err < tol:
            raise SystemExit("parity failure: device %r vs oracle %r (rel %.3e, tol %.0e)"
                             % (dev, ref64, err, tol))

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
            "config": workload_config(a, name, rec, macs),
            "run": {"slices_per_step": a.slices, "slices_per_gpu": len(mine),
                    "kernel_launches_per_slice": sc.program.launches,
                    "arena_bytes": sc.program.arena_bytes, "parallelism": "slices/%d" % world,
                    "slices_in_flight_per_gpu": a.lanes, "zgemm_ozaki": a.ozaki,
                    "ozaki_auto": 0 if a.no_int8 else 1, "host_cpu": host_cpu_model()},
            "amplitude_wall_ms": ms_per_step,
            "amplitude": [float(np.real(amplitude)), float(np.imag(amplitude))],
            "complex_mac_per_s": macs * a.slices / (ms_per_step * 1e-3),
            "tflops_per_gpu": value / world,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s * 1e3},
            "hoisted": hoisted,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "kernels": kernels,
            "peaks": peaks,
            "cpu_baseline": cpu,
        }
        emit(line)
    barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
