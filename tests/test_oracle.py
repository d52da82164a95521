"""Pins the CPU oracle (oracle/qsim_oracle.c) to the reference:
  * against tests/golden/kat_reference.npz, produced by oracle/make_golden.py
    from the unmodified reference SimulatorBasic<float|double>, SimulatorAVX512
    and StateSpaceBasic (always);
  * against oracle/_ref/libqsim_ref_*.so directly, when it has been built.
No GPU needed."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle.make_golden import kat_cases, kat_matrix, kat_state
from oracle.oracle import BASIC_F32, BASIC_F64, SIMD_F32, RefEngine, ref_library_path

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "kat_reference.npz")
KINDS = [("f32", np.complex64, 1e-7), ("f64", np.complex128, 1e-15), ("simd", np.complex64, 2e-7)]


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def host_random_values(num, seed, max_value):
    from qsim_b200 import _lib
    out = np.empty(num, dtype=np.float64)
    rc = _lib.load().qb200_generate_random_values(num, seed, max_value, out.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == 0
    return out


@pytest.mark.parametrize("tag,cdt,tol", KINDS)
def test_gates_and_expectation_match_reference_golden(oracle, golden, tag, cdt, tol):
    for idx, (n, qs, cqs, cvals) in enumerate(kat_cases()):
        st = kat_state(n, cdt, idx)
        m = kat_matrix(len(qs), idx, cdt)
        if len(qs) >= 1 and not cqs:
            ev = oracle.expectation_value(st, qs, m)
            want = golden[f"{tag}/ev/{idx}"]
            assert abs(ev - complex(want[0], want[1])) <= 20 * tol, (idx, qs)
        if cqs:
            oracle.apply_controlled_gate(st, qs, cqs, cvals, m)
        else:
            oracle.apply_gate(st, qs, m)
        assert np.abs(st - golden[f"{tag}/gate/{idx}"]).max() <= tol, (idx, n, qs, cqs, cvals)


@pytest.mark.parametrize("tag,cdt,tol", KINDS)
def test_statespace_matches_reference_golden(oracle, golden, tag, cdt, tol):
    for idx, n in enumerate((1, 3, 8, 12)):
        a, b = kat_state(n, cdt, 500 + idx), kat_state(n, cdt, 600 + idx)
        want = golden[f"{tag}/ss/{idx}/norm_ip"]
        assert abs(oracle.norm(a) - want[0]) < 1e-6
        ip = oracle.inner_product(a, b)
        assert abs(ip - complex(want[1], want[2])) < 1e-6 and abs(ip.real - want[3]) < 1e-6
        # Sample(state, 256, seed=7): host RNG restated in libqsim_b200, scan in the oracle
        rs = host_random_values(256, 7, oracle.sample_norm(a))
        assert np.array_equal(oracle.sample(a, rs), golden[f"{tag}/ss/{idx}/samples"])
        # Measure({0, n-1}) with mt19937(3): r = first uniform(0, norm) draw
        ok, mask, bits = (int(x) for x in golden[f"{tag}/ss/{idx}/measure"])
        assert ok == 1
        r = host_random_values(1, 3, oracle.norm(a))[0]
        assert oracle.find_measured_bits(a, r, mask) == bits
        oracle.collapse(a, mask, bits)
        assert np.abs(a - golden[f"{tag}/ss/{idx}/collapsed"]).max() <= 4 * tol
        oracle.multiply(0.625, b)
        oracle.add(a, b)
        oracle.bulk_set_ampl(b, 1, 1, 0.25 - 0.5j, False)
        assert np.abs(b - golden[f"{tag}/ss/{idx}/mul_add_bulk"]).max() <= 4 * tol


def test_host_rng_is_the_reference_sequence(golden):
    """qb200_generate_random_values == GenerateRandomValues<double> (lib/util.h:67-85)."""
    assert np.array_equal(host_random_values(64, 1, 1.0), golden["rng/seed1_norm1"])
    assert np.array_equal(host_random_values(64, 7, 0.9), golden["rng/seed7_norm0.9"])


@pytest.mark.skipif(ref_library_path() is None, reason="oracle/_ref not built (reference tree absent)")
@pytest.mark.parametrize("kind,cdt,tol", [(BASIC_F32, np.complex64, 1e-7), (BASIC_F64, np.complex128, 1e-15),
                                          (SIMD_F32, np.complex64, 2e-7)])
def test_oracle_matches_live_reference(oracle, kind, cdt, tol):
    """larger randomized comparison against the reference library itself."""
    rs = np.random.RandomState(42)
    n = 15
    st = kat_state(n, cdt, 77)
    e = RefEngine(kind, n, 2)
    e.from_numpy(st)
    for step in range(24):
        g = int(rs.randint(0, 7))
        qs = sorted(rs.choice(n, g, replace=False).tolist())
        m = kat_matrix(g, step, cdt) * (1 << g) ** 0.5
        if step % 3 == 2 and g <= 4:
            free = [q for q in range(n) if q not in qs]
            cqs = sorted(rs.choice(free, int(rs.randint(1, 4)), replace=False).tolist())
            cvals = int(rs.randint(0, 1 << len(cqs)))
            e.apply_controlled_gate(qs, cqs, cvals, m)
            oracle.apply_controlled_gate(st, qs, cqs, cvals, m)
        else:
            e.apply_gate(qs, m)
            oracle.apply_gate(st, qs, m)
        if g >= 1:
            assert abs(e.expectation_value(qs, m) - oracle.expectation_value(st, qs, m)) <= 1e3 * tol * max(1.0, e.norm())
    ref = e.to_numpy()
    scale = np.abs(ref).max()
    assert np.abs(ref - st).max() <= 50 * tol * scale
