"""Trajectory farm (BASELINE config 5): repetition ids are block-partitioned over
GPUs and the per-slice observable sums added -- host logic checked on CPU with the reference-CPU
build of the same driver (oracle/_ref/qsim_qtrajectory_ref) as the worker; on the GPU the B200
worker must reproduce the reference worker's sums for the same seeds."""
import os
import random

import numpy as np
import pytest

from qsim_b200 import traj_farm

REF = os.path.join(traj_farm.ROOT, "oracle", "_ref", "qsim_qtrajectory_ref")


def rotation_circuit(n, depth, seed):
    """qsim-format circuit with generic rotations, so single-qubit expectations are non-trivial."""
    rng = random.Random(seed)
    lines = [str(n)]
    t = 0
    for d in range(depth):
        for q in range(n):
            lines.append(f"{t} {rng.choice(['rx', 'ry', 'rz'])} {q} {rng.uniform(0.2, 2.9):.6f}")
        t += 1
        for a in range(d % 2, n - 1, 2):
            lines.append(f"{t} {rng.choice(['cz', 'is', 'cnot'])} {a} {a + 1}")
        t += 1
    return "\n".join(lines) + "\n"


@pytest.fixture()
def circuit_file(tmp_path):
    p = tmp_path / "rot14"
    p.write_text(rotation_circuit(14, 5, 7))
    return str(p)


def test_partition_covers_ids_once():
    for traj0, num, parts in ((0, 8192, 8), (5, 10, 4), (0, 3, 8), (7, 0, 2)):
        sl = traj_farm.partition(traj0, num, parts)
        assert len(sl) == parts and sum(c for _, c in sl) == num
        ids = [i for s, c in sl for i in range(s, s + c)]
        assert ids == list(range(traj0, traj0 + num))
        assert max(c for _, c in sl) - min(c for _, c in sl) <= 1


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref not built")
def test_slices_add_up(circuit_file):
    one = traj_farm.run_farm(circuit_file, 3, 12, gpus=1, p=0.05, binary=REF, device_ids=[None])
    four = traj_farm.run_farm(circuit_file, 3, 12, gpus=4, p=0.05, binary=REF, device_ids=[None] * 4)
    assert four["gpus"] == 4 and four["num"] == 12
    assert four["gate_passes"] == one["gate_passes"] and four["expect_passes"] == one["expect_passes"]
    assert np.abs(np.array(one["sums"]) - np.array(four["sums"])).max() < 1e-4
    # the observables are not trivially zero for this circuit, and the noise matters
    clean = traj_farm.run_farm(circuit_file, 3, 12, gpus=1, p=0.0, binary=REF, device_ids=[None])
    assert np.abs(np.array(clean["mean"])).max() > 0.1
    assert np.abs(np.array(clean["mean"]) - np.array(one["mean"])).max() > 1e-3


@pytest.mark.gpu
def test_b200_trajectories_match_reference_cpu(circuit_file):
    """same repetition ids => same Kraus choices => same observable sums (fp32 round-off apart)."""
    if not (os.path.exists(REF) and os.path.exists(traj_farm.BINARY)):
        pytest.skip("oracle/_ref or apps/_bin not built (the reference tree was absent at build time)")
    for fused in (2, 4):
        ref = traj_farm.run_farm(circuit_file, 0, 16, gpus=1, p=0.05, max_fused_size=fused, binary=REF,
                                 device_ids=[None])
        # default (-b 2): single-qubit observables from the reduced density matrices (csrc/moments.cu), the
        # Pauli strings batched -- fewer passes, same sums to fp32 round-off
        got = traj_farm.run_farm(circuit_file, 0, 16, gpus=1, p=0.05, max_fused_size=fused)
        # ... and the noiseless prefix of the fused gate list is shared between trajectories (-x 1, default):
        # fewer gate passes, bit-identical sums
        assert got["gate_passes"] + got["prefix_gates_skipped"] == ref["gate_passes"]
        assert got["expect_passes"] < ref["expect_passes"]
        plain = traj_farm.run_farm(circuit_file, 0, 16, gpus=1, p=0.05, max_fused_size=fused, extra_args=("-x", "0"))
        assert plain["gate_passes"] == ref["gate_passes"] and plain["prefix_gates_skipped"] == 0
        assert plain["sums"] == got["sums"]
        err = np.abs(np.array(got["sums"]) - np.array(ref["sums"])).max()
        assert err < 16 * 2e-5, err
        assert np.abs(np.array(ref["mean"])).max() > 0.1
        # "-b 1" = every operator string through its own pass, all enqueued and read after one synchronisation;
        # "-b 0" = the reference's lib/expect.h, one synchronisation per string: same kernels, identical sums
        batched = traj_farm.run_farm(circuit_file, 0, 16, gpus=1, p=0.05, max_fused_size=fused,
                                     extra_args=("-b", "1", "-x", "0"))
        serial = traj_farm.run_farm(circuit_file, 0, 16, gpus=1, p=0.05, max_fused_size=fused,
                                    extra_args=("-b", "0", "-x", "0"))
        assert serial["sums"] == batched["sums"]
        assert serial["expect_passes"] == batched["expect_passes"] == ref["expect_passes"]
        assert np.abs(np.array(got["sums"]) - np.array(serial["sums"])).max() < 16 * 2e-5
        # "-j 3": three worker threads (own state, own CUDA per-thread stream) over sub-slices of the ids:
        # same trajectories, sums added in a different order
        threaded = traj_farm.run_farm(circuit_file, 0, 16, gpus=1, p=0.05, max_fused_size=fused,
                                      extra_args=("-j", "3"))
        assert threaded["gate_passes"] + threaded["prefix_gates_skipped"] == ref["gate_passes"] and threaded["num"] == 16
        assert np.abs(np.array(threaded["sums"]) - np.array(got["sums"])).max() < 1e-9


@pytest.mark.gpu
def test_b200_trajectories_match_reference_cpu_at_26_qubits(tmp_path):
    """BASELINE config 5's size (26 qubits, 512 MiB state, depolarizing noise) with a rotation circuit whose
    single-qubit observables are O(0.1) -- on the RQC of the benchmark they are ~1e-10 and a wrong observable
    would go unnoticed: three repetition ids on the B200 backend against the reference-CPU build of the same
    driver (oracle/_ref/qsim_qtrajectory_ref, the reference's own CPU simulator) on the same seeds."""
    if not (os.path.exists(REF) and os.path.exists(traj_farm.BINARY)):
        pytest.skip("oracle/_ref or apps/_bin not built (the reference tree was absent at build time)")
    p = tmp_path / "rot26"
    p.write_text(rotation_circuit(26, 4, 11))
    threads = ("-t", str(os.cpu_count() or 4))
    ref = traj_farm.run_farm(str(p), 5, 3, gpus=1, p=0.02, max_fused_size=4, binary=REF, device_ids=[None],
                             extra_args=threads)
    got = traj_farm.run_farm(str(p), 5, 3, gpus=1, p=0.02, max_fused_size=4)
    serial = traj_farm.run_farm(str(p), 5, 3, gpus=1, p=0.02, max_fused_size=4, extra_args=("-b", "0"))
    assert got["n"] == 26 and got["gate_passes"] + got["prefix_gates_skipped"] == ref["gate_passes"]
    r, g, s = np.array(ref["sums"]), np.array(got["sums"]), np.array(serial["sums"])
    assert np.abs(r / 3).max() > 0.1, "observables must not be trivially zero"
    assert np.sum(np.abs(r / 3) > 0.05) >= 10
    assert np.abs(s - r).max() < 3 * 2e-5, np.abs(s - r).max()   # the reference's expect.h loop on our kernels
    assert np.abs(g - r).max() < 3 * 2e-5, np.abs(g - r).max()   # moments + batched Pauli strings
    # the noise was sampled: at least one trajectory differs from the noiseless one
    clean = traj_farm.run_farm(str(p), 5, 1, gpus=1, p=0.0, max_fused_size=4)
    one = traj_farm.run_farm(str(p), 5, 3, gpus=1, p=0.02, max_fused_size=4)
    assert np.abs(np.array(one["sums"]) / 3 - np.array(clean["sums"])).max() > 1e-3


@pytest.mark.gpu
def test_weak_noise_shares_most_of_the_circuit(circuit_file):
    """p = 0.001 (the benchmark's noise level): most trajectories are noiseless, the others start from a deep
    checkpoint; sums bit-identical to the plain runner and equal to the reference CPU run."""
    if not (os.path.exists(REF) and os.path.exists(traj_farm.BINARY)):
        pytest.skip("oracle/_ref or apps/_bin not built")
    ref = traj_farm.run_farm(circuit_file, 0, 64, gpus=1, p=0.001, binary=REF, device_ids=[None])
    got = traj_farm.run_farm(circuit_file, 0, 64, gpus=1, p=0.001)
    plain = traj_farm.run_farm(circuit_file, 0, 64, gpus=1, p=0.001, extra_args=("-x", "0"))
    assert got["sums"] == plain["sums"]
    assert np.abs(np.array(got["sums"]) - np.array(ref["sums"])).max() < 64 * 2e-5
    assert got["noiseless_trajectories"] >= 32 and got["gate_passes"] < 0.4 * plain["gate_passes"]


@pytest.mark.gpu
def test_amplitude_damping_trajectories_match_reference_cpu(circuit_file):
    """a NON-unitary channel (lib/channels_cirq.h:274-311): every channel flushes the deferred gates and samples
    its Kraus operator from expectation values of K^dagger K on the current state, then renormalises
    (lib/qtrajectory.h:335-372) -- the ExpectationValue / Norm / Multiply path of the backend inside the
    trajectory loop.  Same repetition ids as the reference CPU simulator => same choices => same sums."""
    if not (os.path.exists(REF) and os.path.exists(traj_farm.BINARY)):
        pytest.skip("oracle/_ref or apps/_bin not built")
    args = ("-C", "amplitude_damp")
    ref = traj_farm.run_farm(circuit_file, 0, 6, gpus=1, p=0.05, binary=REF, device_ids=[None], extra_args=args)
    got = traj_farm.run_farm(circuit_file, 0, 6, gpus=1, p=0.05, extra_args=args)
    assert got["gate_passes"] == ref["gate_passes"]
    assert np.abs(np.array(got["sums"]) - np.array(ref["sums"])).max() < 6 * 5e-5
    clean = traj_farm.run_farm(circuit_file, 0, 1, gpus=1, p=0.0, binary=REF, device_ids=[None])
    assert np.abs(np.array(ref["sums"]) / 6 - np.array(clean["sums"])).max() > 1e-3   # the damping matters
    # operator groups (default, -k 1): when the sampling loop gets as far as the second Kraus operator, its probability
    # comes out of the pass that evaluated the first (qb200_expectation_values_multi).  Strong damping so that it does.
    strong = [traj_farm.run_farm(circuit_file, 0, 6, gpus=1, p=0.3, extra_args=args + ("-k", k)) for k in ("1", "0")]
    sref = traj_farm.run_farm(circuit_file, 0, 6, gpus=1, p=0.3, binary=REF, device_ids=[None], extra_args=args)
    assert strong[0]["kraus_group_hits"] > 0 and strong[1]["kraus_group_hits"] == 0
    assert strong[0]["expect_passes"] == strong[1]["expect_passes"]   # same sampling path: same ExpectationValue calls
    assert strong[0]["gate_passes"] == strong[1]["gate_passes"] == sref["gate_passes"]
    for got in strong:
        assert np.abs(np.array(got["sums"]) - np.array(sref["sums"])).max() < 6 * 5e-5
