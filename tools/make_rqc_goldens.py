#!/usr/bin/env python
"""Golden amplitudes for the sharded benchmark circuits (bench.py --gpus N parity block).

rqc_q31 / q32 / q33 (tests/golden/rqc_q<n>_d20_f4.trace: tools/gen_rqc.py circuits fused by the reference fuser)
are run on ONE GPU through the single-GPU path -- the path the reference's own gtest suites and the q24 / q30
full-state comparisons verify -- and 64 amplitudes at fixed indices plus the norm are written to
tests/golden/rqc_amplitudes.json.  rqc_q31 is additionally run on the reference's own AVX-512 CPU simulator
(oracle/_ref) when present, and the two must agree.  Run on the GPU box:
    python tools/make_rqc_goldens.py gpurun_out/rqc_amplitudes.json
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def golden_indices(n, count=64):
    """fixed, spread over all shards and low/high bits"""
    rng = np.random.RandomState(1000 + n)
    idx = [0, 1, 2, 7, (1 << n) - 1, 1 << (n - 1), (1 << (n - 1)) + 5, (1 << (n - 2)) | 3]
    while len(idx) < count:
        idx.append(int(rng.randint(0, 1 << 30)) | (int(rng.randint(0, 1 << (n - 30))) << 30))
    return idx[:count]


def main():
    import qsim_b200
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "rqc_amplitudes.json")
    sizes = [int(x) for x in os.environ.get("QB200_GOLDEN_SIZES", "31,32,33").split(",")]
    res = {}
    for n in sizes:
        nq, ops = qsim_b200.read_trace(os.path.join(ROOT, "tests", "golden", f"rqc_q{n}_d20_f4.trace"))
        assert nq == n
        ss, sim = qsim_b200.StateSpaceB200(np.float32), qsim_b200.SimulatorB200(np.float32)
        st = ss.Create(n)
        assert not ss.IsNull(st), "not enough device memory"
        ss.SetStateZero(st)
        for op in ops:
            if op.controls:
                sim.ApplyControlledGate(op.qubits, op.controls, op.cvals, op.matrix, st)
            else:
                sim.ApplyGate(op.qubits, op.matrix, st)
        idx = golden_indices(n)
        amps = [ss.GetAmpl(st, i) for i in idx]
        entry = {"indices": idx, "amplitudes": [[a.real, a.imag] for a in amps], "norm": ss.Norm(st),
                 "source": "single-GPU path of libqsim_b200 (k_gate_* kernels), one B200", "passes": len(ops)}
        del st
        if n == 31:
            from oracle.oracle import SIMD_F32, RefEngine, ref_library_path
            if ref_library_path():
                ref = RefEngine(SIMD_F32, n, os.cpu_count() or 1)
                ref.set_zero()
                for op in ops:
                    ref.apply_gate(op.qubits, op.matrix)
                ramps = [ref.get_ampl(i) for i in idx]
                err = max(abs(a - b) for a, b in zip(amps, ramps))
                entry["reference_avx512_max_abs_err"] = err
                entry["reference_avx512_amplitudes"] = [[a.real, a.imag] for a in ramps]
                assert err < 1e-7, err
                del ref
        res[f"rqc_q{n}_d20_f4"] = entry
        print(n, "norm", entry["norm"], "amp0", entry["amplitudes"][0], entry.get("reference_avx512_max_abs_err"), flush=True)
    with open(out_path, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
