"""The C-ABI library loads and exports every symbol include/qsim_b200.h declares;
host-only entry points behave; compute entry points fail loudly without a GPU
(no CPU fallback).  No GPU needed."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "qsim_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qb200_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from qsim_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 35
    assert set(syms) == set(_lib.SIGNATURES), set(syms) ^ set(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    from qsim_b200 import _lib
    lib = _lib.load()
    for s in declared_symbols():
        assert hasattr(lib, s), s
    assert lib.qb200_abi_version() == 3


def test_host_only_entry_points():
    from qsim_b200 import _lib
    lib = _lib.load()
    assert lib.qb200_min_size(0) == 2 and lib.qb200_min_size(30) == 2 << 30
    assert lib.qb200_partial_norms_count(5) == 1
    assert lib.qb200_partial_norms_count(13) == 1
    assert lib.qb200_partial_norms_count(20) == 1 << 7
    out = np.empty(8)
    assert lib.qb200_generate_random_values(8, 1, 1.0, out.ctypes.data_as(C.POINTER(C.c_double))) == 0
    assert np.all(np.diff(out) >= 0) and 0 <= out[0] and out[-1] < 1


def test_no_cpu_fallback_without_gpu():
    from qsim_b200 import _lib
    lib = _lib.load()
    n = C.c_int(0)
    has_gpu = lib.qb200_device_count(C.byref(n)) == 0 and n.value > 0
    if has_gpu:
        pytest.skip("GPU present")
    ctx = C.c_void_p()
    assert lib.qb200_ctx_create(-1, C.byref(ctx)) == _lib.ERR_CUDA
    import qsim_b200
    with pytest.raises(qsim_b200.QB200Error):
        qsim_b200.SimulatorB200()


def test_product_never_imports_the_oracle():
    """Only tests/, bench.py and __graft_entry__.py may touch oracle/."""
    for base in ("qsim_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".h", ".cuh", ".cu", ".cc", ".hpp")):
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    assert "oracle" not in text.lower() or "no cpu fallback" in text.lower() or f == "trace.py", os.path.join(dirpath, f)


def test_trace_fixtures_parse():
    import qsim_b200
    n, ops = qsim_b200.read_trace(os.path.join(ROOT, "tests", "golden", "q30_d20_f4.trace"))
    assert n == 30 and len(ops) == 41
    assert sorted(len(o.qubits) for o in ops).count(4) == 36
    assert all(o.matrix.size == 2 << (2 * len(o.qubits)) for o in ops)
    n, ops = qsim_b200.read_trace(os.path.join(ROOT, "tests", "golden", "q24_f4.trace"))
    assert n == 24 and len(ops) == 175
