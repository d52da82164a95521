import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import ctypes as C
        from qsim_b200 import _lib
        n = C.c_int(0)
        return _lib.load().qb200_device_count(C.byref(n)) == 0 and n.value > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device here (GPU tests run under gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


def random_state(n, cdtype, seed):
    rng = np.random.default_rng(seed)
    st = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    st /= np.linalg.norm(st)
    return st.astype(cdtype)


def random_unitary(g, seed, cdtype=np.complex128):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((1 << g, 1 << g)) + 1j * rng.standard_normal((1 << g, 1 << g))
    q, r = np.linalg.qr(a)
    q = q * (np.diag(r) / np.abs(np.diag(r)))
    return q.astype(cdtype)


def random_matrix(g, seed, cdtype=np.complex128):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((1 << g, 1 << g)) + 1j * rng.standard_normal((1 << g, 1 << g))
    return (a / (1 << g)).astype(cdtype)
