"""The reference's pybind layer compiled on the B200 backend (pybind/pybind_main_b200.cpp ->
qsim_b200/qsim_b200_py*.so): circuits are built through the module's own
add_gate / add_matrix_gate / control_last_gate entry points (what qsimcirq calls,
qsimcirq/qsim_circuit.py) and the full state / amplitudes / samples are compared with the oracle."""
import glob
import importlib.util
import os

import numpy as np
import pytest

from conftest import random_unitary

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


_LOADED = {}


def load_module(name="qsim_b200_py"):
    if name in _LOADED:
        return _LOADED[name]
    hits = glob.glob(os.path.join(ROOT, "qsim_b200", name + ".*.so"))
    if not hits:
        pytest.skip("pybind extension not built (needs the reference tree at build time)")
    spec = importlib.util.spec_from_file_location(name, hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _LOADED[name] = mod
    return mod


# qsim_b200_py: single-GPU backend; qsim_b200_sharded_py: state sharded over "gnd" GPUs through B200Runner
# (pybind/pybind_main_b200_sharded.cpp) -- with one GPU in the box the four shards share it
MODULES = [("qsim_b200_py", {}), ("qsim_b200_sharded_py", {"gnd": 4})]


H = (np.array([[1, 1], [1, -1]]) / np.sqrt(2)).astype(np.complex64)
CZ = np.diag([1, 1, 1, -1]).astype(np.complex64)
BE2 = [0, 2, 1, 3]


def build(q, n, oracle):
    """random circuit through the pybind entry points + the same gates on the oracle."""
    rng = np.random.default_rng(7)
    c = q.Circuit()
    c.num_qubits = n
    want = np.zeros(1 << n, np.complex64)
    want[0] = 1
    t = 0
    for a in range(n):
        q.add_gate(q.GateKind.kH, t, [a], {}, c)
        oracle.apply_gate(want, [a], H)
    for layer in range(6):
        t += 1
        perm = rng.permutation(n).tolist()
        for j in range(0, n - 1, 2):
            qs = sorted(perm[j:j + 2])
            if layer % 2 == 0:
                u = random_unitary(2, 100 * layer + j, np.complex64)
                q.add_matrix_gate(t, qs, np.ascontiguousarray(u).view(np.float32).ravel().tolist(), c)
                # Cirq matrices are big-endian in the gate's qubits (gates_cirq.h: MatrixGate2, q0 = MSB);
                # the engine's convention is bit k of the matrix index <-> qs[k]
                oracle.apply_gate(want, qs, np.ascontiguousarray(u[BE2][:, BE2]))
            else:
                q.add_gate(q.GateKind.kCZ, t, qs, {}, c)
                oracle.apply_gate(want, qs, CZ)
    t += 1
    u = random_unitary(1, 5, np.complex64)
    q.add_matrix_gate(t, [3], np.ascontiguousarray(u).view(np.float32).ravel().tolist(), c)
    q.control_last_gate([0, 5], [1, 0], c)
    oracle.apply_controlled_gate(want, [3], [0, 5], 0b01, u)
    return c, want


def options(c, **kw):
    o = {"c": c, "i": "", "z": 0, "f": 4, "v": 0, "s": 1, "t": 1, "r": 1, "gsst": 512, "gdb": 16}
    o.update(kw)
    return o


@pytest.mark.parametrize("module,extra", MODULES)
def test_fullstate_amplitudes_and_samples(oracle, module, extra):
    q0 = load_module()          # circuit classes + builders live in the base module, as with qsimcirq
    q = load_module(module)
    n = 14
    c, want = build(q0, n, oracle)
    opt = lambda c, **kw: options(c, **dict(extra, **kw))  # noqa: E731 -- the module's own options on every call
    for f in (2, 4):
        got = np.asarray(q.qsim_simulate_fullstate(opt(c, f=f), 0)).view(np.complex64)
        assert got.shape == want.shape
        assert np.abs(got - want).max() < 2e-6
    # amplitudes of chosen bitstrings (qsim_simulate): the string's first character is qubit 0
    idx = [0, 1, 5, (1 << n) - 1, 12345 % (1 << n)]
    strings = "\n".join("".join("1" if (i >> b) & 1 else "0" for b in range(n)) for i in idx)
    amps = np.asarray(q.qsim_simulate(opt(c, i=strings)))
    assert np.abs(amps - want[idx]).max() < 2e-6
    # an initial state handed in as a vector, and one handed in as a basis-state index
    init = np.zeros(2 << n, np.float32)
    init[2 * 3] = 1.0
    a = np.asarray(q.qsim_simulate_fullstate(opt(c), init)).view(np.complex64)
    b = np.asarray(q.qsim_simulate_fullstate(opt(c), 3)).view(np.complex64)
    assert np.abs(a - b).max() < 1e-6 and abs(np.vdot(a, a).real - 1) < 1e-5


@pytest.mark.parametrize("module,extra", MODULES)
def test_expectation_values(oracle, module, extra):
    q0 = load_module()
    q = load_module(module)
    n = 10
    c, want = build(q0, n, oracle)
    opt = lambda c, **kw: options(c, **dict(extra, **kw))  # noqa: E731
    s = q0.OpString()
    s.weight = 1.0
    q0.add_gate_to_opstring(q0.GateKind.kZ, [2], s)
    q0.add_gate_to_opstring(q0.GateKind.kX, [7], s)
    got = q.qsim_simulate_expectation_values(opt(c), [([s], 2)], 0)[0]
    z = np.array([[1, 0], [0, -1]], np.complex64)
    x = np.array([[0, 1], [1, 0]], np.complex64)
    ket = want.copy()
    oracle.apply_gate(ket, [2], z)
    oracle.apply_gate(ket, [7], x)
    assert abs(got - np.vdot(want, ket)) < 1e-5
