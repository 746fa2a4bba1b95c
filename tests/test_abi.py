"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and
exports exactly the symbols include/pq_b200.h declares; the integer lowering code
passes its host emulation; and the product path refuses to run without a GPU."""
import os
import re
import subprocess

import numpy as np
import pytest

import picoquant_jl_b200  # noqa: F401
from picoquant_jl_b200.host import b200_backend

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "picoquant.jl_b200", "csrc")


@pytest.fixture(scope="module")
def built():
    subprocess.run(["make", "-C", CSRC, "-j", "8", "all"], check=True, capture_output=True)
    return True


def test_header_and_library_export_the_same_symbols(built):
    with open(os.path.join(ROOT, "include", "pq_b200.h")) as f:
        header = f.read()
    declared = set(re.findall(r"\b(pq_[a-z0-9_]+)\s*\(", header))
    assert declared == set(b200_backend.ABI_SYMBOLS)
    lib = b200_backend.load_library()
    for name in sorted(declared):
        assert getattr(lib, name) is not None
    out = subprocess.run(["nm", "-D", "--defined-only", b200_backend.LIB_PATH],
                         check=True, capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (pq_[a-z0-9_]+)\b", out))
    assert declared <= exported
    assert lib.pq_version().decode().startswith("pq_b200")


def test_lowering_host_emulation(built):
    """tile parameters, swizzle, gather maps and TTGT layouts (csrc/test_lower.cpp)."""
    r = subprocess.run([os.path.join(CSRC, "test_lower")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ALL OK" in r.stdout


def test_no_cpu_fallback(built):
    """Without a CUDA device the backend must fail loudly, never compute on the CPU."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(b200_backend.B200Error):
        b200_backend.B200Backend(np.complex128)


def test_product_path_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "picoquant.jl_b200")
    for base, _, files in os.walk(pkg):
        for name in files:
            if name.endswith((".py", ".cu", ".cpp", ".h")):
                with open(os.path.join(base, name)) as f:
                    text = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), name


def _build_abi_smoke():
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    subprocess.run(["make", "-C", here, "abi_smoke"], check=True, capture_output=True)
    return os.path.join(here, "abi_smoke")


def test_pure_c_driver_builds_and_links():
    """tests/abi_smoke.c (C99, no Python / C++ in the caller) compiles against include/pq_b200.h
    and links libpq_b200.so; without a GPU it must stop at pq_create with exit code 2."""
    import subprocess
    exe = _build_abi_smoke()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode in (0, 2), (r.returncode, r.stdout, r.stderr)


@pytest.mark.gpu
def test_pure_c_driver_runs_on_gpu():
    """The foreign-host call sequence (create -> save -> save_tensors -> contract -> info ->
    save_output -> load -> permute -> delete) from plain C, checked against a scalar loop."""
    import subprocess
    exe = _build_abi_smoke()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "abi_smoke ok" in r.stdout, (r.returncode, r.stdout, r.stderr)
