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


APPS = os.path.join(ROOT, "apps", "_bin")


def run_suite(name, timeout=900, env=None):
    exe = os.path.join(REF, name)
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (reference tree was absent at build time)")
    p = subprocess.run([exe], capture_output=True, text=True, timeout=timeout, env=dict(os.environ, **(env or {})))
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


# ---- the same reference suites on a SHARDED state (include/qsim_b200/*_sharded.h, run_b200.h) ----------------
# QB200_TEST_SHARDS shards; on a one-GPU box they share the device, so the exchange kernels still run.
@pytest.mark.parametrize("shards,swap_mode", [(4, -1), (2, 0), (8, -1)])
def test_reference_simulator_suite_on_a_sharded_state(shards, swap_mode):
    assert run_suite("simulator_b200_sharded_test.x", env={"QB200_TEST_SHARDS": str(shards),
                                                          "QB200_TEST_SWAP_MODE": str(swap_mode)}) == 22


@pytest.mark.parametrize("shards,swap_mode", [(4, -1), (2, 0), (8, -1)])
def test_reference_statespace_suite_on_a_sharded_state(shards, swap_mode):
    # sampling, measurement, collapse, inner products, bulk-set ... (tests/statespace_testfixture.h)
    assert run_suite("statespace_b200_sharded_test.x", env={"QB200_TEST_SHARDS": str(shards),
                                                           "QB200_TEST_SWAP_MODE": str(swap_mode)}) == 26


def test_reference_qtrajectory_suite_on_a_sharded_state_through_b200runner():
    assert run_suite("qtrajectory_b200_sharded_test.x", env={"QB200_TEST_SHARDS": "4"}) == 7


def test_b200runner_known_answers():
    # 4 tests x {single GPU, sharded}: tests/run_qsim_test.cc's known answers through B200Runner
    assert run_suite("run_b200_test.x", env={"QB200_TEST_SHARDS": "4"}) == 8


def parse_amps(stdout):
    amps = {}
    for line in stdout.splitlines():
        m = re.match(r"([01]{3}):\s+(\S+)\s+(\S+)\s+(\S+)", line)
        if m:
            amps[m.group(1)] = complex(float(m.group(2)), float(m.group(3)))
    return amps


Q24_WANT = {"000": complex(1.0311284e-4, 7.1349914e-6), "001": complex(9.1424146e-5, 2.9970953e-4),
            "010": complex(-1.1130853e-4, 4.4225984e-5), "111": complex(-3.8646715e-5, 2.8088354e-4)}


@pytest.mark.parametrize("shards", [1, 4])
def test_qsim_base_cli_q24_matches_reference_output(shards, tmp_path):
    """apps/qsim_base_b200 (reference parser + fuser, B200Runner) on circuit_q24 -f 4: the amplitudes printed by
    the reference's own AVX-512 build, on one GPU and on a 4-shard state; the dumped
    state (-o) loads back (-i) to the same amplitudes."""
    exe = os.path.join(APPS, "qsim_base_b200")
    circ = os.path.join(REF, "circuits", "circuit_q24")
    if not (os.path.exists(exe) and os.path.exists(circ)):
        pytest.skip("qsim_base_b200 / circuit file not built")
    dump = str(tmp_path / "state.f32")
    p = subprocess.run([exe, "-c", circ, "-f", "4", "-v", "0", "-g", str(shards), "-o", dump],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    amps = parse_amps(p.stdout)
    for k, v in Q24_WANT.items():
        assert abs(amps[k] - v) < 2e-8, (k, amps.get(k))
    assert os.path.getsize(dump) == 8 << 24
    # resume from the dump with an identity circuit: same amplitudes
    ident = tmp_path / "identity_q24"
    ident.write_text("24\n0 id1 0\n")
    p2 = subprocess.run([exe, "-c", str(ident), "-g", str(shards), "-i", dump], capture_output=True, text=True, timeout=600)
    assert p2.returncode == 0, p2.stderr
    amps2 = parse_amps(p2.stdout)
    for k in amps:
        assert abs(amps2[k] - amps[k]) < 1e-12


@pytest.mark.parametrize("shards", [1, 2])
def test_qsim_amplitudes_cli_matches_the_reference_app(shards, tmp_path):
    """apps/qsim_amplitudes_b200 against the reference's apps/qsim_amplitudes.cc on its CPU simulator
    (oracle/_ref/qsim_amplitudes_ref): circuit_q24 to depth 14, amplitudes of circuits/bitstrings_q24_s1."""
    exe = os.path.join(APPS, "qsim_amplitudes_b200")
    ref = os.path.join(REF, "qsim_amplitudes_ref")
    circ = os.path.join(REF, "circuits", "circuit_q24")
    bits = os.path.join(REF, "circuits", "bitstrings_q24_s1")
    if not all(os.path.exists(x) for x in (exe, ref, circ, bits)):
        pytest.skip("apps / checker / input files not built")
    ours, theirs = str(tmp_path / "ours.txt"), str(tmp_path / "theirs.txt")
    p = subprocess.run([exe, "-c", circ, "-d", "14", "-i", bits, "-o", ours, "-f", "4", "-g", str(shards)],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    r = subprocess.run([ref, "-c", circ, "-d", "14", "-i", bits, "-o", theirs, "-f", "4", "-t", str(os.cpu_count() or 4)],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr
    import numpy as np
    a, b = np.loadtxt(ours), np.loadtxt(theirs)
    assert a.shape == b.shape and a.shape[0] >= 100
    assert np.abs(a - b).max() < 1e-6
