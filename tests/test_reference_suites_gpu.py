"""Drop-in proof: the REFERENCE's own gtest suites (tests/simulator_testfixture.h,
statespace_testfixture.h, qtrajectory_testfixture.h, hybrid_testfixture.h) and its
qsim_base CLI flow, compiled in place against include/qsim_b200/*.h by
tests/cpp/Makefile (g++ only), run here on the GPU.  The binaries live in
oracle/_ref/ (built in the container where /root/reference exists; they travel to
the GPU box with the snapshot)."""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def run_suite(name, timeout=900):
    exe = os.path.join(REF, name)
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (reference tree was absent at build time)")
    p = subprocess.run([exe], capture_output=True, text=True, timeout=timeout)
    tail = "\n".join(p.stdout.splitlines()[-25:])
    assert p.returncode == 0, tail + p.stderr[-2000:]
    m = re.search(r"\[  PASSED  \] (\d+) tests", p.stdout)
    assert m, tail
    return int(m.group(1))


def test_reference_simulator_suite():
    # 11 known-answer tests x {float, double} (tests/simulator_cuda_test.cu:50-126)
    assert run_suite("simulator_b200_test.x") == 22


def test_reference_statespace_suite():
    # 12 typed tests x {float, double} + MeasurementSmall + InvalidStateSize
    assert run_suite("statespace_b200_test.x") == 26


def test_reference_qtrajectory_suite():
    assert run_suite("qtrajectory_b200_test.x") == 7


def test_reference_hybrid_suite():
    assert run_suite("hybrid_b200_test.x") == 2


def test_qsim_base_cli_q24_matches_reference_output():
    """apps/qsim_base.cc flow (parser + fuser + QSimRunner unchanged) on circuit_q24 -f 4:
    amplitudes printed by the reference's own AVX-512 build."""
    exe = os.path.join(REF, "qsim_base_b200")
    circ = os.path.join(REF, "circuits", "circuit_q24")
    if not (os.path.exists(exe) and os.path.exists(circ)):
        pytest.skip("qsim_base_b200 / circuit file not built")
    p = subprocess.run([exe, "-c", circ, "-f", "4", "-v", "0"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    amps = {}
    for line in p.stdout.splitlines():
        m = re.match(r"([01]{3}):\s+(\S+)\s+(\S+)\s+(\S+)", line)
        if m:
            amps[m.group(1)] = complex(float(m.group(2)), float(m.group(3)))
    want = {"000": complex(1.0311284e-4, 7.1349914e-6), "001": complex(9.1424146e-5, 2.9970953e-4),
            "010": complex(-1.1130853e-4, 4.4225984e-5), "111": complex(-3.8646715e-5, 2.8088354e-4)}
    for k, v in want.items():
        assert abs(amps[k] - v) < 2e-8, (k, amps.get(k))
