"""Parity of the CUDA gate path (through the C ABI) against the CPU oracle.

Mirrors the reference's own simulator tests (tests/simulator_testfixture.h):
TestMultiQubitGates (:1147-1212), TestControlledGates (:1215-1362),
TestGlobalPhaseGate (:1365-1413), TestExpectationValue1 (:1416-1472).
Tolerances: fp32 per-amplitude |d| <= 1e-5 (north_star), in practice ~1e-7;
fp64 <= 1e-12.
"""
import itertools

import numpy as np
import pytest

from conftest import random_matrix, random_state, random_unitary

pytestmark = pytest.mark.gpu

TOL = {np.complex64: 2e-6, np.complex128: 1e-13}
RDT = {np.complex64: np.float32, np.complex128: np.float64}


def backends(cdt):
    import qsim_b200
    return qsim_b200.StateSpaceB200(RDT[cdt]), qsim_b200.SimulatorB200(RDT[cdt])


def target_sets(n, g):
    """lowest, highest, scattered and mixed target layouts."""
    if g == 0:
        return [[]]
    sets = {tuple(range(g)), tuple(range(n - g, n))}
    if n >= 2 * g:
        sets.add(tuple(range(0, 2 * g, 2)))
        sets.add(tuple(range(1, 2 * g, 2)))
    if n >= 6 + g:
        sets.add(tuple(range(5, 5 + g)))
        sets.add(tuple([0] + list(range(6, 4 + g)) + [n - 1]) if g >= 2 else (n - 1,))
    rng = np.random.default_rng(100 * n + g)
    for _ in range(2):
        sets.add(tuple(sorted(rng.choice(n, size=g, replace=False).tolist())))
    return [list(s) for s in sets if len(s) == g]


@pytest.mark.parametrize("cdt", [np.complex64, np.complex128])
@pytest.mark.parametrize("n", [1, 2, 3, 5, 6, 9, 13, 16])
def test_apply_gate_matches_oracle(oracle, cdt, n):
    ss, sim = backends(cdt)
    for g in range(0, min(n, 6) + 1):
        for k, qs in enumerate(target_sets(n, g)):
            host = random_state(n, cdt, seed=n * 1000 + g * 10 + k)
            m = random_unitary(g, seed=g * 7 + k, cdtype=cdt)
            st = ss.Create(n)
            ss.from_numpy(host, st)
            sim.ApplyGate(qs, m, st)
            got = ss.to_numpy(st)
            want = oracle.apply_gate(host.copy(), qs, m)
            err = np.abs(got - want).max()
            assert err <= TOL[cdt], (n, qs, err)


@pytest.mark.parametrize("cdt", [np.complex64, np.complex128])
def test_kernel_variants_agree(oracle, cdt):
    """register kernels (1 and 2 amplitudes per thread) and the runtime-generic
    kernel must all reproduce the oracle."""
    n = 14
    for variant in ({"gate_mode": 0}, {"force_generic": 1}, {"tile": 0, "tc": 0}, {"tile": 1, "tc": 0}, {"tile": 2, "tc": 0},
                    {"tile": 0, "prefetch": 0, "tc": 0}, {"tile": 1, "gate_mode": 0, "tc": 0}, {"tc": 0}, {"tc": 0, "big": 0},
                    {"tc_low": 0}):
        ss, sim = backends(cdt)
        for key, val in variant.items():
            sim.set_tuning(key, val)
        for g in range(0, 7):
            for k, qs in enumerate(target_sets(n, g)):
                host = random_state(n, cdt, seed=g * 10 + k)
                m = random_matrix(g, seed=g * 7 + k, cdtype=cdt)
                st = ss.Create(n)
                ss.from_numpy(host, st)
                sim.ApplyGate(qs, m, st)
                err = np.abs(ss.to_numpy(st) - oracle.apply_gate(host.copy(), qs, m)).max()
                assert err <= TOL[cdt], (variant, qs, err)


@pytest.mark.parametrize("cdt", [np.complex64, np.complex128])
@pytest.mark.parametrize("n", [3, 6, 8, 11])
def test_controlled_gates_match_oracle(oracle, cdt, n):
    """exhaustive-ish control/target masks, cvals all-0 / all-1 / mixed, non-unitary
    matrix (tests/simulator_testfixture.h:1215-1362)."""
    ss, sim = backends(cdt)
    rng = np.random.default_rng(n)
    cases = []
    for g in range(0, min(4, n - 1) + 1):
        for c in range(1, min(3, n - g) + 1):
            for _ in range(4):
                perm = rng.permutation(n)[: g + c].tolist()
                qs, cqs = sorted(perm[:g]), sorted(perm[g:])
                for cvals in {0, (1 << c) - 1, int(rng.integers(0, 1 << c))}:
                    cases.append((qs, cqs, cvals))
    # low controls and low targets explicitly
    if n >= 6:
        cases += [([1, 3], [0, 2], 0b10), ([0], [1, 2, 3], 0b111), ([5], [0], 1), ([0, 1, 2, 3], [4], 0),
                  ([], [0], 1), ([], [2, 5], 0b01)]
    for k, (qs, cqs, cvals) in enumerate(cases):
        host = random_state(n, cdt, seed=k)
        m = random_matrix(len(qs), seed=k, cdtype=cdt)
        st = ss.Create(n)
        ss.from_numpy(host, st)
        sim.ApplyControlledGate(qs, cqs, cvals, m, st)
        got = ss.to_numpy(st)
        want = oracle.apply_controlled_gate(host.copy(), qs, cqs, cvals, m)
        err = np.abs(got - want).max()
        assert err <= TOL[cdt], (qs, cqs, cvals, err)


@pytest.mark.parametrize("cdt", [np.complex64, np.complex128])
@pytest.mark.parametrize("n", [1, 4, 7, 12, 16])
def test_expectation_value_matches_oracle(oracle, cdt, n):
    ss, sim = backends(cdt)
    for g in range(1, min(n, 6) + 1):
        for k, qs in enumerate(target_sets(n, g)):
            host = random_state(n, cdt, seed=g * 10 + k)
            m = random_matrix(g, seed=g + k, cdtype=cdt)
            st = ss.Create(n)
            ss.from_numpy(host, st)
            got = sim.ExpectationValue(qs, m, st)
            want = oracle.expectation_value(host, qs, m)
            tol = 1e-6 if cdt == np.complex64 else 1e-13
            assert abs(got - want) <= tol, (n, qs, got, want)
            # read-only: state unchanged, bit for bit
            assert np.array_equal(ss.to_numpy(st), host)


@pytest.mark.parametrize("cdt", [np.complex64, np.complex128])
def test_batched_expectation_values_equal_single_calls(oracle, cdt):
    """qb200_reduce_batch_begin/end (include/qsim_b200/expect_b200.h): the passes of a batch are the
    same kernels in the same order, read after one synchronisation -> bit-identical values; the
    batch also survives growing past its pre-sized slot array and leaves the context usable."""
    ss, sim = backends(cdt)
    n = 14
    host = random_state(n, cdt, seed=5)
    st = ss.Create(n)
    ss.from_numpy(host, st)
    terms = []
    for rep in range(3):  # 3 x 60 terms > the 64 slots allocated first
        for g in range(1, 7):
            for k, qs in enumerate(target_sets(n, g)):
                terms.append((qs, random_matrix(g, seed=100 * rep + g + k, cdtype=cdt)))
    single = [sim.ExpectationValue(qs, m, st) for qs, m in terms]
    batch = sim.ExpectationValues(terms[:7], st) + sim.ExpectationValues(terms[7:], st)
    assert len(batch) == len(terms)
    assert batch == single
    want = oracle.expectation_value(host, terms[0][0], terms[0][1])
    assert abs(batch[0] - want) <= (1e-6 if cdt == np.complex64 else 1e-13)
    assert sim.ExpectationValues([], st) == []
    # a batch cannot be opened twice, and ending without begin is an error, not a hang
    assert sim._lib.qb200_reduce_batch_end(sim._ctx, None, 0, None) != 0
    assert sim.ExpectationValue(terms[0][0], terms[0][1], st) == single[0]
    assert np.array_equal(ss.to_numpy(st), host)


@pytest.mark.parametrize("cdt", [np.complex64, np.complex128])
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 8, 11, 12, 13, 15, 17, 20, 21])
def test_one_qubit_moments(oracle, cdt, n):
    """qb200_one_qubit_moments (csrc/moments.cu): S00, S11, S01 of every qubit against numpy in double, and
    <M_q> rebuilt from them against the oracle's ExpectationValue (lib/simulator_basic.h:296-342) and the
    per-operator kernel; tile layouts: one partial tile (n <= 12 / 11), several passes with filler bits."""
    ss, sim = backends(cdt)
    host = random_state(n, cdt, seed=50 + n)
    st = ss.Create(n)
    ss.from_numpy(host, st)
    got = sim.OneQubitMoments(st)
    assert got.shape == (n, 4)
    psi = host.astype(np.complex128)
    tol = 2e-6 if cdt == np.complex64 else 1e-13
    for q in range(n):
        v = psi.reshape(1 << (n - 1 - q), 2, 1 << q)
        a0, a1 = v[:, 0, :].ravel(), v[:, 1, :].ravel()
        s01 = np.vdot(a0, a1)
        want = [np.vdot(a0, a0).real, np.vdot(a1, a1).real, s01.real, s01.imag]
        assert np.abs(got[q] - np.array(want)).max() <= tol, (n, q, got[q], want)
    for q in sorted({0, n // 2, n - 1}):
        m = random_matrix(1, seed=q, cdtype=cdt)
        rebuilt = sim.moment_expectation(got[q], m)
        assert abs(rebuilt - oracle.expectation_value(host, [q], m)) <= tol * 4
        assert abs(rebuilt - sim.ExpectationValue([q], m, st)) <= tol * 4
    assert np.array_equal(ss.to_numpy(st), host)  # read-only
    assert np.array_equal(sim.OneQubitMoments(st), got)  # deterministic


_PAULI = {
    "I": np.eye(2), "X": np.array([[0, 1], [1, 0]]), "Y": np.array([[0, -1j], [1j, 0]]), "Z": np.diag([1, -1]),
    "S": np.diag([1, 1j]), "T": np.diag([1, np.exp(0.25j * np.pi)]), "P0": np.diag([1, 0]),  # projector: zero row
    "XS": np.array([[0, 1], [1j, 0]]), "H": np.array([[1, 1], [1, -1]]) / np.sqrt(2),        # H: not monomial
}


def _string_matrix(names, weight, cdt):
    """kron with names[k] on matrix-index bit k <-> qs[k] (lib/matrix.h:26-33)."""
    m = np.array([[weight]], dtype=np.complex128)
    for name in names:
        m = np.kron(_PAULI[name], m)
    return m.astype(cdt)


@pytest.mark.parametrize("cdt", [np.complex64, np.complex128])
@pytest.mark.parametrize("n", [6, 9, 14, 18])
def test_pauli_string_expectation_read_pass(oracle, cdt, n):
    """XOR-monomial operators (Pauli strings, phase gates, projectors) on 3..6 targets take the read pass of
    csrc/expect_monomial.cu; it must agree with the oracle (lib/simulator_basic.h:296-342) and with the dense
    kernels (tuning mono = 0), and anything with a Hadamard in it must fall through to the dense kernels."""
    ss, sim = backends(cdt)
    _, dense = backends(cdt)
    dense.set_tuning("mono", 0)
    host = random_state(n, cdt, seed=70 + n)
    st = ss.Create(n)
    ss.from_numpy(host, st)
    rng = np.random.default_rng(n)
    tol = 1e-6 if cdt == np.complex64 else 1e-13
    cases = [(list("XYZ"), None), (list("ZZZ"), None), (list("IIII"), None), (list("XZYXZY"), None),
             (list("ZZZZZZ"), None), (list("YYYYY"), None), (list("SXTZ"), None), (["P0", "X", "Z"], None),
             (["XS", "XS", "Y"], None), (list("XHZ"), None), (list("ZZHZZZ"), None),
             (list("XXX"), [0, 1, 2]), (list("XZYXZY"), list(range(6))), (list("YZX"), [n - 3, n - 2, n - 1]),
             (list("IZX"), [0, 3, n - 1])]
    for names, qs in cases:
        g = len(names)
        if qs is None:
            qs = sorted(rng.choice(n, size=g, replace=False).tolist())
        assert len(qs) == g
        m = _string_matrix(names, 0.7 - 0.4j, cdt)
        got = sim.ExpectationValue(qs, m, st)
        want = oracle.expectation_value(host, qs, m)
        assert abs(got - want) <= tol, (names, qs, got, want)
        assert abs(got - dense.ExpectationValue(qs, m, st)) <= 2 * tol, (names, qs)
    # batched: the read pass fills result slots like every other reduction
    terms = [(list(range(6)), _string_matrix(list("XZYXZY"), 1.0, cdt)), ([1, 2, 4], _string_matrix(list("ZZZ"), 1.0, cdt))]
    assert sim.ExpectationValues(terms, st) == [sim.ExpectationValue(q, m, st) for q, m in terms]
    assert np.array_equal(ss.to_numpy(st), host)


def test_gate_application_is_deterministic():
    """EXPECT_EQ bit-identical amplitudes across repeated runs
    (tests/simulator_testfixture.h:735-764)."""
    import qsim_b200
    ss, sim = qsim_b200.StateSpaceB200(np.float32), qsim_b200.SimulatorB200(np.float32)
    n = 18
    host = random_state(n, np.complex64, 3)
    outs = []
    for _ in range(3):
        st = ss.Create(n)
        ss.from_numpy(host, st)
        for k, qs in enumerate([[0, 1, 2, 3], [4, 9, 13, 17], [2, 6], [1, 5, 7, 11, 16]]):
            sim.ApplyGate(qs, random_unitary(len(qs), k, np.complex64), st)
        sim.ApplyControlledGate([3, 8], [0, 12], 0b11, random_unitary(2, 9, np.complex64), st)
        outs.append(ss.to_numpy(st))
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])


def test_unsupported_sizes_are_ignored():
    """>6 targets, or >4 targets under control: state untouched (lib/simulator_cuda.h:96-98,162-164)."""
    import qsim_b200
    ss, sim = qsim_b200.StateSpaceB200(np.float32), qsim_b200.SimulatorB200(np.float32)
    n = 9
    host = random_state(n, np.complex64, 5)
    st = ss.Create(n)
    ss.from_numpy(host, st)
    sim.ApplyGate(list(range(7)), np.zeros(2 << 14, np.float32), st)
    sim.ApplyControlledGate([0, 1, 2, 3, 4], [8], 1, np.zeros(2 << 10, np.float32), st)
    assert np.array_equal(ss.to_numpy(st), host)


def test_unitarity_round_trip_large():
    """size-independent property at a BASELINE-like size: U then U^dagger restores the state."""
    import qsim_b200
    ss, sim = qsim_b200.StateSpaceB200(np.float32), qsim_b200.SimulatorB200(np.float32)
    n = 26
    st = ss.Create(n)
    ss.SetStateUniform(st)
    layouts = [[0, 1, 2, 3], [3, 9, 17, 25], [22, 23, 24, 25], [5, 6], [0, 25], [1, 8, 13, 19, 24], [2, 4, 6, 10, 20, 21]]
    us = [random_unitary(len(q), i, np.complex64) for i, q in enumerate(layouts)]
    for q, u in zip(layouts, us):
        sim.ApplyGate(q, u, st)
    assert abs(ss.Norm(st) - 1.0) < 1e-5
    for q, u in reversed(list(zip(layouts, us))):
        sim.ApplyGate(q, np.ascontiguousarray(u.conj().T), st)
    amp = 2.0 ** (-n / 2)
    got = ss.to_numpy(st)
    assert np.abs(got - amp).max() < 1e-5 * amp * 100
    assert abs(ss.Norm(st) - 1.0) < 1e-5


def test_tile_kernel_every_4_qubit_layout(oracle):
    """warp-cooperative swizzled-tile kernel (gate_tile.cuh): EVERY choice of 4 targets out of 11
    qubits (330 layouts: all low/high mixes, every swizzle pattern), plus controlled variants."""
    import qsim_b200
    ss, sim = qsim_b200.StateSpaceB200(np.float32), qsim_b200.SimulatorB200(np.float32)
    sim.set_tuning("tile", 2)
    sim.set_tuning("tc", 0)  # keep the 4-qubit gates off the tensor-core kernel
    n = 11
    host = random_state(n, np.complex64, seed=77)
    st = ss.Create(n)
    worst = 0.0
    for k, qs in enumerate(itertools.combinations(range(n), 4)):
        m = random_matrix(4, seed=k, cdtype=np.complex64)
        ss.from_numpy(host, st)
        sim.ApplyGate(list(qs), m, st)
        err = np.abs(ss.to_numpy(st) - oracle.apply_gate(host.copy(), list(qs), m)).max()
        worst = max(worst, err)
        assert err <= 2e-6, (qs, err)
    rng = np.random.default_rng(5)
    for k in range(60):
        perm = rng.permutation(n)
        qs, cqs = sorted(perm[:4].tolist()), sorted(perm[4:4 + int(rng.integers(1, 3))].tolist())
        cvals = int(rng.integers(0, 1 << len(cqs)))
        m = random_matrix(4, seed=1000 + k, cdtype=np.complex64)
        ss.from_numpy(host, st)
        sim.ApplyControlledGate(qs, cqs, cvals, m, st)
        err = np.abs(ss.to_numpy(st) - oracle.apply_controlled_gate(host.copy(), qs, cqs, cvals, m)).max()
        assert err <= 2e-6, (qs, cqs, cvals, err)


@pytest.mark.parametrize("g", [5, 6])
def test_big_kernel_every_layout(oracle, g):
    """row-blocked warp-tile kernel for 5- and 6-qubit gates (gate_big.cuh): EVERY choice of g
    targets out of 12 qubits (792 / 924 layouts) for ApplyGate, and every third layout for
    ExpectationValue; both launch shapes."""
    import qsim_b200
    ss, sim = qsim_b200.StateSpaceB200(np.float32), qsim_b200.SimulatorB200(np.float32)
    sim.set_tuning("tc", 0)  # 5/6-qubit gates and expectations default to the tensor-core kernels;
    sim.set_tuning("tcx", 0)  # this test is about k_gate_big
    n = 12
    host = random_state(n, np.complex64, seed=g)
    st = ss.Create(n)
    for k, qs in enumerate(itertools.combinations(range(n), g)):
        sim.set_tuning("big", 1 if k % 2 else -1)
        m = random_matrix(g, seed=k % 17, cdtype=np.complex64)
        ss.from_numpy(host, st)
        sim.ApplyGate(list(qs), m, st)
        err = np.abs(ss.to_numpy(st) - oracle.apply_gate(host.copy(), list(qs), m)).max()
        assert err <= 4e-6, (qs, err)
        if k % 3 == 0:
            ss.from_numpy(host, st)
            got = sim.ExpectationValue(list(qs), m, st)
            want = oracle.expectation_value(host, list(qs), m)
            assert abs(got - want) <= 2e-6 * (1 << g), (qs, got, want)
            assert np.array_equal(ss.to_numpy(st), host)


def test_big_kernel_agrees_with_generic(oracle):
    """tuning big=0 routes 5/6-qubit gates through the register/generic kernels: same answer."""
    import qsim_b200
    ss, sim = qsim_b200.StateSpaceB200(np.float32), qsim_b200.SimulatorB200(np.float32)
    n = 15
    host = random_state(n, np.complex64, seed=2)
    for qs in ([0, 1, 2, 3, 4], [1, 4, 6, 9, 14], [0, 2, 3, 5, 8, 13], [9, 10, 11, 12, 13, 14]):
        m = random_unitary(len(qs), seed=len(qs), cdtype=np.complex64)
        outs = []
        sim.set_tuning("tc", 0)
        for big in (-1, 0):
            sim.set_tuning("big", big)
            st = ss.Create(n)
            ss.from_numpy(host, st)
            sim.ApplyGate(qs, m, st)
            outs.append(ss.to_numpy(st))
        assert np.abs(outs[0] - outs[1]).max() <= 2e-6, qs
        assert np.abs(outs[0] - oracle.apply_gate(host.copy(), qs, m)).max() <= 2e-6, qs


@pytest.mark.parametrize("g", [4, 5])
def test_tensor_core_kernel_every_layout(oracle, g):
    """tcgen05 3xTF32 kernels (gate_tc.cuh): EVERY choice of g targets out of 12 qubits (495 / 792
    layouts: 8-byte and 16-byte row pieces, every swizzle phase), cycling through the variants
    (tuning tc = 1, 2: operands through shared memory; 3: A operand in tensor memory + bias
    compensation = the default; 4: no compensation; 5: four CTAs per SM), plus controlled variants
    for g = 4.  Same tolerance as the fp32 CUDA-core kernels: the 3xTF32 split keeps fp32-level
    accuracy (measured max |d| 2e-8 on normalised states)."""
    import qsim_b200
    ss, sim = qsim_b200.StateSpaceB200(np.float32), qsim_b200.SimulatorB200(np.float32)
    n = 12
    host = random_state(n, np.complex64, seed=40 + g)
    st = ss.Create(n)
    for k, qs in enumerate(itertools.combinations(range(n), g)):
        sim.set_tuning("tc", 1 + (k % 5))
        m = random_matrix(g, seed=k % 13, cdtype=np.complex64)
        ss.from_numpy(host, st)
        sim.ApplyGate(list(qs), m, st)
        err = np.abs(ss.to_numpy(st) - oracle.apply_gate(host.copy(), list(qs), m)).max()
        assert err <= 2e-6, (qs, err)
    if g == 4:
        rng = np.random.default_rng(11)
        for k in range(40):
            perm = rng.permutation(n)
            qs, cqs = sorted(perm[:4].tolist()), sorted(perm[4:5].tolist())
            cvals = int(rng.integers(0, 2))
            m = random_matrix(4, seed=500 + k, cdtype=np.complex64)
            sim.set_tuning("tc", 3 if k % 2 else 1)
            ss.from_numpy(host, st)
            sim.ApplyControlledGate(qs, cqs, cvals, m, st)
            err = np.abs(ss.to_numpy(st) - oracle.apply_controlled_gate(host.copy(), qs, cqs, cvals, m)).max()
            assert err <= 2e-6, (qs, cqs, cvals, err)


@pytest.mark.parametrize("variant,drift", [(1, 5e-7), (3, 1e-7)])
def test_tensor_core_kernel_large_state_round_trip(variant, drift):
    """size-independent property at n = 26 through many persistent tiles per CTA: U then U^dagger on the
    tensor-core path restores the state; norm drift per pass stays below 5e-7 for plain 3xTF32 and
    well below that with the accumulation-bias compensation of the default variant (the six passes
    include CUDA-core layouts and the fp32 reduction of Norm itself)."""
    import qsim_b200
    ss, sim = qsim_b200.StateSpaceB200(np.float32), qsim_b200.SimulatorB200(np.float32)
    sim.set_tuning("tc", variant)
    n = 26
    st = ss.Create(n)
    ss.SetStateUniform(st)
    layouts = [[3, 9, 17, 25], [0, 5, 11, 20], [1, 8, 13, 19, 24], [22, 23, 24, 25], [0, 1, 2, 3, 4], [6, 7, 8, 9, 10]]
    us = [random_unitary(len(q), i, np.complex64) for i, q in enumerate(layouts)]
    for q, u in zip(layouts, us):
        sim.ApplyGate(q, u, st)
    assert abs(ss.Norm(st) - 1.0) < len(layouts) * drift
    for q, u in reversed(list(zip(layouts, us))):
        sim.ApplyGate(q, np.ascontiguousarray(u.conj().T), st)
    amp = 2.0 ** (-n / 2)
    assert np.abs(ss.to_numpy(st) - amp).max() < 1e-5 * amp * 100


def test_tensor_core_6_qubit_gates_and_expectations(oracle):
    """k_gate_tcx (gate_tc.cuh): 6-qubit gates and 4/5/6-qubit expectation values on the tensor cores.
    Every third choice of 6 targets out of 13 qubits for the gate (572 layouts), every fifth choice of
    4 / 5 / 6 targets for the expectation value, all against the oracle; the FFMA2 kernels (tuning
    tcx = 0) must agree too."""
    import qsim_b200
    ss, sim = qsim_b200.StateSpaceB200(np.float32), qsim_b200.SimulatorB200(np.float32)
    n = 13
    host = random_state(n, np.complex64, seed=66)
    st = ss.Create(n)
    for k, qs in enumerate(itertools.combinations(range(n), 6)):
        if k % 3:
            continue
        m = random_matrix(6, seed=k % 7, cdtype=np.complex64)
        ss.from_numpy(host, st)
        sim.ApplyGate(list(qs), m, st)
        err = np.abs(ss.to_numpy(st) - oracle.apply_gate(host.copy(), list(qs), m)).max()
        assert err <= 4e-6, (qs, err)
    ss.from_numpy(host, st)
    for g in (4, 5, 6):
        for k, qs in enumerate(itertools.combinations(range(n), g)):
            if k % 5:
                continue
            m = random_matrix(g, seed=k % 11, cdtype=np.complex64)
            want = oracle.expectation_value(host, list(qs), m)
            sim.set_tuning("tcx", -1)
            got = sim.ExpectationValue(list(qs), m, st)
            assert abs(got - want) <= 2e-6 * (1 << g), (qs, got, want)
            if k % 50 == 0:
                sim.set_tuning("tcx", 0)
                assert abs(sim.ExpectationValue(list(qs), m, st) - want) <= 2e-6 * (1 << g), qs
    assert np.array_equal(ss.to_numpy(st), host)  # expectation values are read-only


def test_tensor_core_6_qubit_gate_norm_drift():
    """the compensation constant of the 6-qubit tensor-core gate: 16 random unitaries at n = 24 keep the norm"""
    import qsim_b200
    ss, sim = qsim_b200.StateSpaceB200(np.float32), qsim_b200.SimulatorB200(np.float32)
    n = 24
    st = ss.Create(n)
    ss.SetStateUniform(st)
    rng = np.random.default_rng(9)
    for i in range(16):
        qs = sorted(rng.choice(np.arange(3, n), 6, replace=False).tolist())
        sim.ApplyGate(qs, random_unitary(6, i, np.complex64), st)
    assert abs(ss.Norm(st) - 1.0) < 16 * 1e-7


def _structured_matrices(g, seed):
    """permutation with phases in {+-1, +-i}, diagonal phases, identity, a Pauli string: one non-zero per row"""
    rng = np.random.default_rng(seed)
    dim = 1 << g
    perm = np.zeros((dim, dim), np.complex64)
    perm[np.arange(dim), rng.permutation(dim)] = rng.choice(np.array([1, -1, 1j, -1j], np.complex64), dim)
    diag = np.diag(np.exp(1j * rng.uniform(0, 2 * np.pi, dim))).astype(np.complex64)
    ident = np.eye(dim, dtype=np.complex64)
    p = {"X": np.array([[0, 1], [1, 0]]), "Y": np.array([[0, -1j], [1j, 0]]), "Z": np.diag([1, -1])}
    pauli = np.array([[1.0]])
    for c in rng.choice(list("XYZ"), g):
        pauli = np.kron(p[c], pauli)
    return {"permutation": perm, "diagonal": diag, "identity": ident, "pauli": pauli.astype(np.complex64)}


@pytest.mark.parametrize("g", [4, 5, 6])
def test_tensor_core_path_on_structured_matrices_and_sparse_states(oracle, g):
    """the accumulation-bias compensation of the tensor-core kernels was fitted on dense unitaries and
    dense states.  Matrices with one non-zero per row get none (gates_f32_tc.cu is_monomial): on basis states
    permutations / Pauli strings / identities leave the state EXACT (bit for bit against the oracle); on dense
    states every amplitude is reproduced to the 22 significant bits the hi + lo TF32 split of the state carries
    (<= 2^-22 relative per pass, no accumulation); diagonal phases (one complex product per amplitude, plain
    3xTF32) drift by less than 1.5e-7 per pass (measured 6e-8: the truncating accumulation of the two real
    products)."""
    import qsim_b200
    ss, sim = qsim_b200.StateSpaceB200(np.float32), qsim_b200.SimulatorB200(np.float32)
    n = 18
    mats = _structured_matrices(g, g)
    rng = np.random.default_rng(100 + g)
    if g <= 5:
        sim.set_tuning("tc", 3)   # the tensor-core kernel for EVERY layout (the default sends a few low-target G = 4 layouts to FFMA2 kernels)
    for start in ("basis", "dense"):
        for name, m in mats.items():
            host = np.zeros(1 << n, np.complex64)
            if start == "basis":
                host[12345] = 1
            else:
                host = random_state(n, np.complex64, seed=5)
            st = ss.Create(n)
            ss.from_numpy(host, st)
            want = host.copy()
            for it in range(100):
                qs = sorted(rng.choice(np.arange(4 if it % 2 else 0, n), g, replace=False).tolist())
                sim.ApplyGate(qs, m, st)
                assert sim.last_kernel_name().startswith("k_gate_tc"), sim.last_kernel_name()
                if name != "diagonal" and it < 12:
                    oracle.apply_gate(want, qs, m)
                    if it == 11:
                        got = ss.to_numpy(st)
                        if start == "basis":
                            assert np.array_equal(got, want), (start, name)
                        else:  # 12 passes, each within 2^-22 of the amplitude's larger component
                            assert np.abs(got - want).max() <= 12 * 2.0 ** -22 * np.abs(want).max(), (start, name)
            assert abs(ss.Norm(st) - 1.0) < (1.5e-5 if name == "diagonal" else 1e-6), (start, name, ss.Norm(st))
    # dense unitaries on a basis state (the first passes of every circuit): the compensated path keeps the norm
    sim.set_tuning("tc", -1)
    st = ss.Create(n)
    ss.SetStateZero(st)
    for it in range(60):
        qs = sorted(rng.choice(np.arange(n), g, replace=False).tolist())
        sim.ApplyGate(qs, random_unitary(g, it, np.complex64), st)
    assert abs(ss.Norm(st) - 1.0) < 60 * 2e-7


@pytest.mark.parametrize("cdt", [np.complex64, np.complex128])
def test_expectation_values_of_several_operators_in_one_pass(oracle, cdt):
    """qb200_expectation_values_multi (csrc/expect_multi.cu): up to 8 operators on the same 1 or 2 qubits, one read
    pass -- each value equals the oracle's ExpectationValue of that operator (lib/simulator_basic.h:286-342
    arithmetic: products in the state's precision, sums in double), for low / high / mixed qubits, non-Hermitian
    matrices included; more than 2 qubits or 8 operators is refused without touching the output."""
    import qsim_b200
    from qsim_b200 import _lib
    rdt = np.float32 if cdt == np.complex64 else np.float64
    ss, sim = qsim_b200.StateSpaceB200(rdt), qsim_b200.SimulatorB200(rdt)
    n = 17
    h = random_state(n, cdt, 77)
    st = ss.Create(n)
    ss.from_numpy(h, st)
    rng = np.random.default_rng(5)
    tol = 2e-6 if cdt == np.complex64 else 1e-13
    for qs in ([0], [1], [4], [n - 1], [0, 1], [0, 9], [3, 4], [7, n - 1], [n - 2, n - 1]):
        d = 1 << len(qs)
        for count in (1, 2, 5, 8):
            ms = [(rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d))).astype(cdt) for _ in range(count)]
            got = sim.ExpectationValuesSameQubits(qs, ms, st)
            for m, v in zip(ms, got):
                assert abs(v - oracle.expectation_value(h, qs, m)) < tol * d * 4, (qs, count)
    assert np.array_equal(ss.to_numpy(st), h)   # read-only
    with pytest.raises(qsim_b200.QB200Error) as e:
        sim.ExpectationValuesSameQubits([0, 1, 2], [np.eye(8)], st)
    assert e.value.status == _lib.ERR_UNSUPPORTED
    with pytest.raises(qsim_b200.QB200Error):
        sim.ExpectationValuesSameQubits([0], [np.eye(2)] * 9, st)


@pytest.mark.parametrize("g", [4, 5, 6])
def test_fp64_register_blocked_gemm_kernel(oracle, g):
    """k_gate_d6 (csrc/gate_dbig.cuh): fp64 gates and expectation values on 4, 5, 6 qubits as a register-blocked
    complex GEMM over tiles of 2^(12 - G) groups -- every class of layout (low / high / mixed targets, target 0
    in and out), a controlled 4-qubit gate, against the oracle in double; the round-1 kernels (tuning big = 3) agree."""
    import qsim_b200
    ss, sim = qsim_b200.StateSpaceB200(np.float64), qsim_b200.SimulatorB200(np.float64)
    n = 15
    h = random_state(n, np.complex128, 40 + g)
    rng = np.random.default_rng(g)
    layouts = [list(range(g)), list(range(n - g, n)), list(range(3, 3 + g)), [0] + list(range(n - g + 1, n)),
               [1] + list(range(6, 5 + g))] + [sorted(rng.choice(n, g, replace=False).tolist()) for _ in range(4)]
    for qs in layouts:
        u = random_unitary(g, 7 * g + qs[0], np.complex128)
        st = ss.Create(n)
        ss.from_numpy(h, st)
        gemm = not (g == 4 and qs[0] == 0)   # 4 targets with bit 0 among them stay on the register kernels
        ev = sim.ExpectationValue(qs, u, st)
        assert (sim.last_kernel_name() == "k_gate_d6") == gemm
        assert abs(ev - oracle.expectation_value(h, qs, u)) < 1e-12
        assert np.array_equal(ss.to_numpy(st), h)
        sim.ApplyGate(qs, u, st)
        assert (sim.last_kernel_name() == "k_gate_d6") == gemm
        want = h.copy()
        oracle.apply_gate(want, qs, u)
        got = ss.to_numpy(st)
        assert np.abs(got - want).max() < 1e-14, qs
        sim.set_tuning("big", 3)
        ss.from_numpy(h, st)
        sim.ApplyGate(qs, u, st)
        assert sim.last_kernel_name() != "k_gate_d6"
        assert np.abs(ss.to_numpy(st) - got).max() < 1e-14
        sim.set_tuning("big", -1)
    if g == 4:
        u = random_unitary(4, 3, np.complex128)
        st = ss.Create(n)
        ss.from_numpy(h, st)
        sim.ApplyControlledGate([2, 5, 9, 12], [0, 14], 0b10, u, st)
        want = h.copy()
        oracle.apply_controlled_gate(want, [2, 5, 9, 12], [0, 14], 0b10, u)
        assert np.abs(ss.to_numpy(st) - want).max() < 1e-14
