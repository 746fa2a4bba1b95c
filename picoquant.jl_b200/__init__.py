"""pq-b200: a B200-native tensor-network contraction backend for PicoQuant.

The package holds only what the contraction hot path needs: ``csrc/`` (the
sm_100a CUDA kernels and the C-ABI shared library ``libpq_b200.so``) and
``host/`` (a Python mirror of the reference's backend interface and of the
layer-2/3 host bookkeeping that drives it).  See DESIGN.md.
"""
__version__ = "0.1.0"
