"""End-to-end parity on the BASELINE circuits: the fused-gate traces written by
the reference's own parser + fuser (tests/golden/*.trace) are replayed through
the C ABI and compared with (a) the CPU oracle over the FULL state and (b) the
amplitudes the reference's qsim_base printed."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def run_trace(name, tuning=None):
    import qsim_b200
    n, ops = qsim_b200.read_trace(os.path.join(GOLDEN, name))
    ss, sim = qsim_b200.StateSpaceB200(np.float32), qsim_b200.SimulatorB200(np.float32)
    for k, v in (tuning or {}).items():
        sim.set_tuning(k, v)
    st = ss.Create(n)
    ss.SetStateZero(st)
    for op in ops:
        if op.controls:
            sim.ApplyControlledGate(op.qubits, op.controls, op.cvals, op.matrix, st)
        else:
            sim.ApplyGate(op.qubits, op.matrix, st)
    return n, ops, ss, st


def test_circuit_q24_depth20_full_state_vs_oracle(oracle):
    n, ops, ss, st = run_trace("q24_d20_f4.trace")
    got = ss.to_numpy(st)
    want = np.zeros(1 << n, np.complex64)
    want[0] = 1
    for op in ops:
        oracle.apply_gate(want, op.qubits, op.matrix)
    assert np.abs(got - want).max() <= 1e-6  # north_star budget is 1e-5
    fid = abs(np.vdot(want.astype(np.complex128), got.astype(np.complex128))) ** 2
    assert fid >= 1 - 1e-5
    assert abs(ss.Norm(st) - 1) < 1e-5


def test_circuit_q24_full_depth_known_amplitudes():
    """circuit_q24, all 100 time steps, -f 4 (175 fused gates): amplitudes printed by
    the reference's apps/qsim_base.cc (AVX-512 build), BASELINE.md section 4."""
    n, ops, ss, st = run_trace("q24_f4.trace")
    assert len(ops) == 175
    want = {0: (1.0311284e-4, 7.1349914e-6), 1: (9.1424146e-5, 2.9970953e-4),
            2: (-1.1130853e-4, 4.4225984e-5), 7: (-3.8646715e-5, 2.8088354e-4)}
    for i, (re, im) in want.items():
        a = ss.GetAmpl(st, i)
        assert abs(a - complex(re, im)) < 2e-8, (i, a)
    assert abs(ss.Norm(st) - 1) < 1e-4


@pytest.mark.parametrize("trace", ["q30_d20_f4.trace", "q30_d20_f5.trace"])
def test_circuit_q30_depth20_known_amplitudes(trace):
    """BASELINE config 2 at full size (8 GiB state): reference qsim_base amplitudes."""
    n, ops, ss, st = run_trace(trace)
    assert n == 30
    want = {0: (1.4871957e-5, 2.8161678e-5), 1: (1.8767701e-5, 7.3190154e-6),
            2: (-1.1130518e-5, 1.6207156e-5), 7: (-1.622395e-5, 3.5199686e-5)}
    for i, (re, im) in want.items():
        a = ss.GetAmpl(st, i)
        assert abs(a - complex(re, im)) < 2e-9, (i, a)
    assert abs(ss.Norm(st) - 1) < 1e-4
    # sampling at full size: sorted draws -> non-decreasing indices, all with non-zero probability
    samples = ss.Sample(st, 1000, 1)
    assert np.all(np.diff(samples.astype(np.int64)) >= 0)
