"""``B200Backend``: PicoQuant's backend interface on top of ``libpq_b200.so``.

The Python twin of ``InteractiveBackend{T}`` (``src/backends/interactive.jl``)
whose tensor store lives in B200 HBM behind the C ABI declared in
``include/pq_b200.h``.  The nine backend functions keep the reference's names,
argument meaning (1-based, column-major) and error behaviour (missing label ->
``KeyError``; ``load_tensor_data`` of a missing label -> ``None``;
``delete_tensor`` of a missing label is fine).

There is **no CPU fallback**: if the shared library is missing, cannot be
loaded, or no CUDA device is usable, construction raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, byref, c_char_p, c_double, c_int, c_int32, c_int64, c_void_p)
from typing import List, Optional, Sequence

import numpy as np

from .backends import AbstractBackend, Metrics

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "csrc", "libpq_b200.so")

PQ_C64, PQ_C128 = 0, 1
PQ_HOST_F32, PQ_HOST_F64, PQ_HOST_C64, PQ_HOST_C128 = 0, 1, 2, 3
PQ_MAX_RANK = 64
PQ_NUM_KERNEL_CLASSES = 15

_STATUS = {-1: "PQ_ERR_INVALID", -2: "PQ_ERR_NOT_FOUND", -3: "PQ_ERR_SHAPE", -4: "PQ_ERR_CUDA",
           -5: "PQ_ERR_NCCL", -6: "PQ_ERR_PARSE", -7: "PQ_ERR_UNSUPPORTED"}

# every symbol include/pq_b200.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "pq_create", "pq_destroy", "pq_last_error", "pq_version", "pq_save_tensor",
    "pq_tensor_info", "pq_load_tensor", "pq_contract", "pq_permute", "pq_reshape", "pq_view",
    "pq_delete", "pq_save_output", "pq_sync", "pq_accumulate", "pq_comm_unique_id",
    "pq_comm_init", "pq_allreduce_sum", "pq_program_compile", "pq_program_num_views",
    "pq_program_run", "pq_program_destroy", "pq_program_stats", "pq_get_counters",
    "pq_reset_counters", "pq_profile_enable", "pq_profile_read", "pq_kernel_class_name",
    "pq_set_option", "pq_microbench", "pq_timer_begin", "pq_timer_end",
    "pq_program_set_hoist", "pq_program_prepare", "pq_program_hoist_stats",
    "pq_program_run_slices", "pq_decompose", "pq_save_tensors", "pq_program_profile_slices",
]

_lib = None


class B200Error(RuntimeError):
    pass


def load_library(path: Optional[str] = None) -> ctypes.CDLL:
    """Loads ``libpq_b200.so`` (built in-tree by ``__graft_entry__.build()``) and
    declares the prototypes.  Raises if the library is absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.isfile(p):
        raise B200Error("libpq_b200.so not found at %s -- run __graft_entry__.build() "
                        "(there is no CPU fallback)" % p)
    lib = ctypes.CDLL(p)
    i32p, i64p = POINTER(c_int32), POINTER(c_int64)
    lib.pq_version.restype = c_char_p
    lib.pq_last_error.restype = c_char_p
    lib.pq_last_error.argtypes = [c_void_p]
    lib.pq_kernel_class_name.restype = c_char_p
    lib.pq_kernel_class_name.argtypes = [c_int]
    lib.pq_create.argtypes = [c_int, c_int, POINTER(c_void_p)]
    lib.pq_destroy.argtypes = [c_void_p]
    lib.pq_save_tensor.argtypes = [c_void_p, c_char_p, c_int, i64p, c_void_p, c_int]
    lib.pq_save_tensors.argtypes = [c_void_p, c_int, POINTER(c_char_p), POINTER(c_int), i64p,
                                    POINTER(c_void_p), POINTER(c_int)]
    lib.pq_tensor_info.argtypes = [c_void_p, c_char_p, POINTER(c_int), i64p]
    lib.pq_load_tensor.argtypes = [c_void_p, c_char_p, c_void_p, c_int]
    lib.pq_contract.argtypes = [c_void_p, c_char_p, i32p, c_int, c_char_p, i32p, c_int, c_char_p]
    lib.pq_permute.argtypes = [c_void_p, c_char_p, i32p, c_int]
    lib.pq_reshape.argtypes = [c_void_p, c_char_p, i32p, i32p, c_int]
    lib.pq_view.argtypes = [c_void_p, c_char_p, c_char_p, c_int, i32p, c_int]
    lib.pq_decompose.argtypes = [c_void_p, c_char_p, i32p, c_int, i32p, c_int, c_double, c_int,
                                 c_char_p, c_char_p, POINTER(c_int)]
    lib.pq_delete.argtypes = [c_void_p, c_char_p]
    lib.pq_save_output.argtypes = [c_void_p, c_char_p, c_char_p]
    lib.pq_sync.argtypes = [c_void_p]
    lib.pq_accumulate.argtypes = [c_void_p, c_char_p, c_char_p]
    lib.pq_comm_unique_id.argtypes = [c_void_p]
    lib.pq_comm_init.argtypes = [c_void_p, c_void_p, c_int, c_int]
    lib.pq_allreduce_sum.argtypes = [c_void_p, c_char_p]
    lib.pq_program_compile.argtypes = [c_void_p, c_char_p, POINTER(c_void_p)]
    lib.pq_program_num_views.argtypes = [c_void_p]
    lib.pq_program_run.argtypes = [c_void_p, c_void_p, i32p, c_int, c_char_p]
    lib.pq_program_run_slices.argtypes = [c_void_p, c_void_p, i32p, c_int, c_int, c_char_p, c_int]
    dp = POINTER(c_double)
    lib.pq_program_profile_slices.argtypes = [c_void_p, c_void_p, i32p, c_int, c_int, c_int, dp, dp,
                                              i64p, dp, dp, dp]
    lib.pq_program_destroy.argtypes = [c_void_p, c_void_p]
    lib.pq_program_stats.argtypes = [c_void_p, i64p, i64p, i64p]
    lib.pq_program_set_hoist.argtypes = [c_void_p, c_int]
    lib.pq_program_prepare.argtypes = [c_void_p, c_void_p]
    lib.pq_program_hoist_stats.argtypes = [c_void_p, i64p, i64p, i64p, i64p]
    lib.pq_get_counters.argtypes = [c_void_p, i64p, i64p, i64p, i64p]
    lib.pq_reset_counters.argtypes = [c_void_p]
    lib.pq_profile_enable.argtypes = [c_void_p, c_int]
    lib.pq_profile_read.argtypes = [c_void_p, POINTER(c_double), i64p, POINTER(c_double),
                                    POINTER(c_double)]
    lib.pq_set_option.argtypes = [c_void_p, c_char_p, c_int]
    lib.pq_microbench.argtypes = [c_void_p, c_char_p, POINTER(c_double)]
    lib.pq_timer_begin.argtypes = [c_void_p]
    lib.pq_timer_end.argtypes = [c_void_p, POINTER(c_double)]
    for name in ABI_SYMBOLS:
        fn = getattr(lib, name)
        if fn.restype is not c_char_p:
            fn.restype = c_int
    if path is None:
        _lib = lib
    return lib


def _i32(values: Sequence[int]):
    arr = (c_int32 * max(1, len(values)))(*[int(v) for v in values])
    return arr


_HOST_DTYPES = {np.dtype(np.float32): PQ_HOST_F32, np.dtype(np.float64): PQ_HOST_F64,
                np.dtype(np.complex64): PQ_HOST_C64, np.dtype(np.complex128): PQ_HOST_C128}


class Program:
    """A compiled ``.tl`` command stream (``pq_program``)."""

    def __init__(self, backend: "B200Backend", text: str) -> None:
        self.backend = backend
        self._p = c_void_p()
        backend._check(backend.lib.pq_program_compile(backend._h, text.encode(), byref(self._p)))
        self.num_views = backend.lib.pq_program_num_views(self._p)
        a, l, m = c_int64(), c_int64(), c_int64()
        backend.lib.pq_program_stats(self._p, byref(a), byref(l), byref(m))
        self.arena_bytes, self.launches, self.macs = a.value, l.value, m.value
        mi, md, li, ld = c_int64(), c_int64(), c_int64(), c_int64()
        backend.lib.pq_program_hoist_stats(self._p, byref(mi), byref(md), byref(li), byref(ld))
        self.macs_invariant, self.macs_dependent = mi.value, md.value
        self.launches_invariant, self.launches_dependent = li.value, ld.value

    def set_hoist(self, on: bool) -> None:
        """Slice-invariant hoisting: ``prepare()`` runs the part of the stream that does not
        depend on any ``view`` index once; ``run()`` then replays only the rest."""
        self.backend._check(self.backend.lib.pq_program_set_hoist(self._p, 1 if on else 0))

    def prepare(self) -> None:
        self.backend._check(self.backend.lib.pq_program_prepare(self.backend._h, self._p))

    def run(self, view_starts: Optional[Sequence[int]] = None,
            accumulate_into: Optional[str] = None) -> None:
        b = self.backend
        if view_starts is None:
            vs, n = None, 0
        else:
            vs, n = _i32(view_starts), len(view_starts)
        acc = accumulate_into.encode() if accumulate_into else None
        b._check(b.lib.pq_program_run(b._h, self._p, vs, n, acc))

    def run_slices(self, view_starts: Sequence[Sequence[int]], accumulate_into: Optional[str],
                   lanes: int = 2) -> None:
        """The slice loop in one call (``pq_program_run_slices``): one run per row of
        ``view_starts``, results accumulated in row order, up to ``lanes`` slices in flight."""
        b = self.backend
        n = len(view_starts)
        nv = len(view_starts[0]) if n else 0
        flat = _i32([v for row in view_starts for v in row])
        acc = accumulate_into.encode() if accumulate_into else None
        b._check(b.lib.pq_program_run_slices(b._h, self._p, flat if nv else None, n, nv, acc,
                                             int(lanes)))

    def profile_slices(self, view_starts: Sequence[Sequence[int]], lanes: int = 1) -> dict:
        """Per-kernel-class timing measured INSIDE the graph replays of the slice loop
        (``pq_program_profile_slices``): same lanes and parallel branches as ``run_slices``,
        with event-record nodes around every kernel node.  Returns ``{"wall_ms": ..,
        "classes": {name: {busy_ms, sum_ms, launches, bytes, flops}}}``; ``busy_ms`` is the
        union of the class's kernel intervals (<= wall_ms)."""
        b = self.backend
        n = len(view_starts)
        nv = len(view_starts[0]) if n else 0
        flat = _i32([v for row in view_starts for v in row])
        k = PQ_NUM_KERNEL_CLASSES
        busy, tot = (c_double * k)(), (c_double * k)()
        la = (c_int64 * k)()
        by, fl = (c_double * k)(), (c_double * k)()
        wall = c_double()
        b._check(b.lib.pq_program_profile_slices(b._h, self._p, flat if nv else None, n, nv,
                                                 int(lanes), busy, tot, la, by, fl, byref(wall)))
        classes = {}
        for i in range(k):
            if la[i]:
                classes[b.lib.pq_kernel_class_name(i).decode()] = {
                    "busy_ms": busy[i], "sum_ms": tot[i], "launches": la[i], "bytes": by[i],
                    "flops": fl[i]}
        return {"wall_ms": wall.value, "classes": classes}

    def close(self) -> None:
        if self._p:
            self.backend.lib.pq_program_destroy(self.backend._h, self._p)
            self._p = c_void_p()

    def __del__(self):
        try:
            if self.backend._h:
                self.close()
        except Exception:
            pass


class B200Backend(AbstractBackend):
    """``InteractiveBackend{CuArray{T}}`` re-thought: labels -> tensors in HBM.

    ``dtype`` is the backend element type (``np.complex64`` default, like the
    reference's default constructor, or ``np.complex128``)."""

    def __init__(self, dtype=np.complex64, device: int = 0) -> None:
        self.lib = load_library()
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.dtype(np.complex64), np.dtype(np.complex128)):
            raise ValueError("backend dtype must be complex64 or complex128")
        self.metrics = Metrics()
        self._h = c_void_p()
        rc = self.lib.pq_create(int(device), PQ_C128 if self.dtype == np.complex128 else PQ_C64,
                                byref(self._h))
        if rc != 0 or not self._h:
            self._h = c_void_p()
            raise B200Error("pq_create failed (%s): a CUDA device is required, there is no CPU "
                            "fallback" % _STATUS.get(rc, rc))
        self.device = device
        # A/B knobs from the environment (profiling runs): PQ_B200_OPTS="graph=1,gemm=1"
        for item in os.environ.get("PQ_B200_OPTS", "").split(","):
            if "=" in item:
                k, v = item.split("=", 1)
                self.set_option(k.strip(), int(v))

    # -- plumbing -----------------------------------------------------------
    def _check(self, rc: int) -> None:
        if rc == 0:
            return
        msg = self.lib.pq_last_error(self._h).decode(errors="replace")
        if rc == -2:
            raise KeyError(msg)
        if rc == -3:
            raise ValueError("DimensionMismatch: " + msg)
        raise B200Error("%s: %s" % (_STATUS.get(rc, rc), msg))

    def close(self) -> None:
        if self._h:
            self.lib.pq_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- the nine backend functions -----------------------------------------
    def save_tensor_data(self, tensor_label, tensor_data):
        arr = np.asarray(tensor_data)
        if arr.dtype not in _HOST_DTYPES:
            arr = arr.astype(np.complex128 if np.iscomplexobj(arr) else np.float64)
        arr = np.asarray(arr, order="F")
        if not arr.flags.f_contiguous:
            arr = np.array(arr, order="F")
        dims = (c_int64 * max(1, arr.ndim))(*arr.shape)
        self._check(self.lib.pq_save_tensor(self._h, tensor_label.encode(), arr.ndim, dims,
                                            arr.ctypes.data_as(c_void_p), _HOST_DTYPES[arr.dtype]))

    def prepare_save(self, tensor_label, tensor_data):
        """Pre-marshals the arguments of ``pq_save_tensor`` for a host array that
        will be uploaded repeatedly (the array is kept alive by the tuple)."""
        arr = np.asarray(tensor_data)
        if arr.dtype not in _HOST_DTYPES:
            arr = arr.astype(np.complex128 if np.iscomplexobj(arr) else np.float64)
        arr = np.array(arr, order="F", copy=True) if not arr.flags.f_contiguous else arr
        dims = (c_int64 * max(1, arr.ndim))(*arr.shape)
        args = (self._h, tensor_label.encode(), arr.ndim, dims, arr.ctypes.data_as(c_void_p),
                _HOST_DTYPES[arr.dtype], arr)
        return args, arr.size * self.dtype.itemsize

    def prepare_save_batch(self, items):
        """Pre-marshals ``pq_save_tensors`` for ``[(label, array), ...]`` (the gate tensors of
        a network, uploaded again for every amplitude): returns ``(args, nbytes)``; the host
        arrays are kept alive by the tuple."""
        arrs, labels = [], []
        for label, data in items:
            arr = np.asarray(data)
            if arr.dtype not in _HOST_DTYPES:
                arr = arr.astype(np.complex128 if np.iscomplexobj(arr) else np.float64)
            arr = np.array(arr, order="F", copy=True) if not arr.flags.f_contiguous else arr
            arrs.append(arr)
            labels.append(label.encode())
        n = len(arrs)
        flat = [d for a in arrs for d in a.shape]
        args = (self._h, n, (c_char_p * max(1, n))(*labels),
                (c_int * max(1, n))(*[a.ndim for a in arrs]),
                (c_int64 * max(1, len(flat)))(*flat),
                (c_void_p * max(1, n))(*[a.ctypes.data for a in arrs]),
                (c_int * max(1, n))(*[_HOST_DTYPES[a.dtype] for a in arrs]), arrs, labels)
        return args, sum(a.size for a in arrs) * self.dtype.itemsize

    def save_prepared_batch(self, args):
        rc = self.lib.pq_save_tensors(*args[:7])
        if rc != 0:
            self._check(rc)

    def save_tensors(self, items):
        """``save_tensor_data`` for many tensors in one call (one staging copy, one H2D,
        one scatter launch): ``items`` = ``[(label, array), ...]``."""
        args, _ = self.prepare_save_batch(list(items))
        self.save_prepared_batch(args)

    def save_prepared(self, args):
        rc = self.lib.pq_save_tensor(*args[:6])
        if rc != 0:
            self._check(rc)

    def tensor_shape(self, tensor_label):
        rank = c_int()
        dims = (c_int64 * PQ_MAX_RANK)()
        rc = self.lib.pq_tensor_info(self._h, tensor_label.encode(), byref(rank), dims)
        if rc == -2:
            return None
        self._check(rc)
        return tuple(dims[i] for i in range(rank.value))

    def load_tensor_data(self, tensor_label):
        shape = self.tensor_shape(tensor_label)
        if shape is None:
            return None
        out = np.empty(shape, dtype=self.dtype, order="F")
        code = PQ_HOST_C128 if self.dtype == np.complex128 else PQ_HOST_C64
        self._check(self.lib.pq_load_tensor(self._h, tensor_label.encode(),
                                            out.ctypes.data_as(c_void_p), code))
        return out

    def contract_tensors(self, A_label, A_ncon_indices, B_label, B_ncon_indices, C_label):
        self._check(self.lib.pq_contract(self._h, A_label.encode(), _i32(A_ncon_indices),
                                         len(A_ncon_indices), B_label.encode(),
                                         _i32(B_ncon_indices), len(B_ncon_indices),
                                         C_label.encode()))

    def save_output(self, node, name="result"):
        self._check(self.lib.pq_save_output(self._h, node.encode(), name.encode()))

    def reshape_tensor(self, tensor, groups):
        flat: List[int] = [int(x) for g in groups for x in g]
        sizes = [len(g) for g in groups]
        self._check(self.lib.pq_reshape(self._h, tensor.encode(), _i32(flat), _i32(sizes),
                                        len(sizes)))

    def permute_tensor(self, tensor, axes):
        self._check(self.lib.pq_permute(self._h, tensor.encode(), _i32(axes), len(axes)))

    def decompose_tensor(self, tensor, left_positions, right_positions, *, threshold=1e-13,
                         max_rank=0, left_label, right_label):
        """``decompose_tensor!`` (interactive.jl:130-152): SVD split on the device
        (``pq_decompose``, one-sided Jacobi); returns the new bond dimension chi."""
        chi = c_int()
        self._check(self.lib.pq_decompose(self._h, tensor.encode(), _i32(left_positions),
                                          len(left_positions), _i32(right_positions),
                                          len(right_positions), float(threshold), int(max_rank),
                                          left_label.encode(), right_label.encode(), byref(chi)))
        return chi.value

    def delete_tensor(self, tensor_label):
        self.lib.pq_delete(self._h, tensor_label.encode())

    def view_tensor(self, view_node, node, bond_idx, bond_range):
        idx = list(bond_range)
        self._check(self.lib.pq_view(self._h, view_node.encode(), node.encode(), int(bond_idx),
                                     _i32(idx), len(idx)))

    # -- beyond the nine: sliced accumulation, programs, instrumentation ------
    def sync(self):
        self._check(self.lib.pq_sync(self._h))

    def accumulate(self, dst, src):
        self._check(self.lib.pq_accumulate(self._h, dst.encode(), src.encode()))

    @staticmethod
    def comm_unique_id() -> bytes:
        lib = load_library()
        buf = ctypes.create_string_buffer(128)
        rc = lib.pq_comm_unique_id(buf)
        if rc != 0:
            raise B200Error("pq_comm_unique_id failed: %s" % _STATUS.get(rc, rc))
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, nranks: int):
        buf = ctypes.create_string_buffer(unique_id, 128)
        self._check(self.lib.pq_comm_init(self._h, buf, rank, nranks))

    def allreduce_sum(self, label):
        self._check(self.lib.pq_allreduce_sum(self._h, label.encode()))

    def compile_program(self, tl_text: str) -> Program:
        return Program(self, tl_text)

    def execute_dsl(self, tl_text: str, store, output_store=None) -> None:
        """``execute_dsl_file(dsl_file, tensor_data_file, output_data_file)``
        (``src/layer1.jl:211-315``) on the device: ``tensor`` commands read from
        ``store`` (the HDF5 stand-in), ``save`` commands write to ``output_store``
        (default: ``store``).  A stream without ``decompose`` is compiled into one
        program (static arena + CUDA graph).  ``decompose`` makes the shapes depend on
        run-time singular values, so such a stream is interpreted command by command
        through the same backend calls instead."""
        from .backends import parse_dsl
        ops = parse_dsl(tl_text)
        out = output_store if output_store is not None else store
        if not any(cmd == "decompose" for cmd, _ in ops):
            for cmd, a in ops:
                if cmd == "tensor":
                    self.save_tensor_data(a["key"], store.read(a["key"]))
            prog = self.compile_program(tl_text)
            try:
                prog.run()
                for cmd, a in ops:
                    if cmd == "save":
                        out.write(a["key"], self.load_tensor_data(a["key"]))
            finally:
                prog.close()
            return
        for cmd, a in ops:
            if cmd == "tensor":
                self.save_tensor_data(a["t"], store.read(a["key"]))
            elif cmd == "ncon":       # consumes A and B; the stream's own `del` lines are no-ops
                self.contract_tensors(a["A"], a["a_idx"], a["B"], a["b_idx"], a["C"])
            elif cmd == "del":
                self.delete_tensor(a["t"])
            elif cmd == "reshape":
                self.reshape_tensor(a["t"], a["groups"])
            elif cmd == "permute":
                self.permute_tensor(a["t"], a["axes"])
            elif cmd == "view":
                self.view_tensor(a["v"], a["t"], a["axis"], a["idx"])
            elif cmd == "decompose":
                self.decompose_tensor(a["t"], a["left_idx"], a["right_idx"],
                                      threshold=float(a["options"].get("threshold", 1e-13)),
                                      max_rank=int(a["options"].get("max_rank", 0)),
                                      left_label=a["left"], right_label=a["right"])
            elif cmd == "save":
                out.write(a["key"], self.load_tensor_data(a["t"]))

    def counters(self):
        a, b, c, d = c_int64(), c_int64(), c_int64(), c_int64()
        self._check(self.lib.pq_get_counters(self._h, byref(a), byref(b), byref(c), byref(d)))
        return {"n_contract": a.value, "macs": b.value, "max_elems": c.value,
                "kernel_launches": d.value}

    def reset_counters(self):
        self.lib.pq_reset_counters(self._h)

    def profile_enable(self, on: bool):
        self._check(self.lib.pq_profile_enable(self._h, 1 if on else 0))

    def profile_read(self):
        n = PQ_NUM_KERNEL_CLASSES
        ms, la = (c_double * n)(), (c_int64 * n)()
        by, fl = (c_double * n)(), (c_double * n)()
        self._check(self.lib.pq_profile_read(self._h, ms, la, by, fl))
        out = {}
        for i in range(n):
            if la[i]:
                out[self.lib.pq_kernel_class_name(i).decode()] = {
                    "ms": ms[i], "launches": la[i], "bytes": by[i], "flops": fl[i]}
        return out

    def timer_begin(self):
        self._check(self.lib.pq_timer_begin(self._h))

    def timer_end(self) -> float:
        """Milliseconds of device time on the handle's stream since ``timer_begin``."""
        ms = c_double()
        self._check(self.lib.pq_timer_end(self._h, byref(ms)))
        return ms.value

    def set_option(self, key: str, value: int):
        self._check(self.lib.pq_set_option(self._h, key.encode(), int(value)))

    def microbench(self, what: str) -> float:
        r = c_double()
        self._check(self.lib.pq_microbench(self._h, what.encode(), byref(r)))
        return r.value
