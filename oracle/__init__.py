"""CPU oracle for the tensor-contraction hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, in NumPy (OpenBLAS), the reference's in-memory CPU
contraction: ``src/layer1.jl`` (array math), ``src/backends/interactive.jl``
(storage semantics) and the interpreter half of ``src/layer1.jl:211-315``
(``execute_dsl_file``).  It exists to check the CUDA path and to be timed as the
CPU baseline; nothing in the product path (``picoquant.jl_b200/``) imports it.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may use it.

Parity status: the reference cannot be executed in this image (no Julia, no
TensorOperations.jl -- un-vendored, compat "3.0.0", no Manifest pin).  The
oracle is therefore pinned against the reference's own fixtures and
known-answer tests: ``examples/ghz_3.json`` + ``ghz_3_plan.json``
-> ``ghz_3_contracted.json`` (bit-for-bit layout and values), the metrics golden
8/44/124, GHZ / H⊗H / QFT-vs-inverse-FFT known answers, the decomposed-gate
re-contraction identity and the slicing identity.  Bit-level output of
``tensorcontract`` itself is pinned by no reference test (all use ``≈``), so
floating-point parity is tolerance-based (1e-10 c128 / 1e-5 c64 rel-L2).
"""
