#!/usr/bin/env python
"""Checks the tcgen05 3xTF32 gate kernels (tuning tc=1) against the CPU oracle at small n and times
them against the CUDA-core kernels at n=30.  usage: tools/tc_check.py [--n 30] [--skip-timing]"""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qsim_b200
from oracle.oracle import Oracle

ap = argparse.ArgumentParser(); ap.add_argument("--n", type=int, default=30); ap.add_argument("--skip-timing", action="store_true")
args = ap.parse_args()
orc = Oracle()
ss, sim = qsim_b200.StateSpaceB200(np.float32), qsim_b200.SimulatorB200(np.float32)
rng = np.random.RandomState(5)

def unitary(g, seed):
    r = np.random.RandomState(seed)
    a = r.standard_normal((1 << g, 1 << g)) + 1j * r.standard_normal((1 << g, 1 << g))
    u, _ = np.linalg.qr(a)
    return u.astype(np.complex64)

worst = {}
for n in (12, 15):
    host = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    host = (host / np.linalg.norm(host)).astype(np.complex64)
    for g in (4, 5):
        layouts = [list(range(g)), list(range(n - g, n)), list(range(3, 3 + g)), [0] + list(range(n - g + 1, n)),
                   sorted(rng.choice(n, g, replace=False).tolist()), sorted(rng.choice(n, g, replace=False).tolist()),
                   [1] + list(range(5, 4 + g)), [2, 3] + list(range(n - g + 2, n))]
        for qs in layouts:
            u = unitary(g, g + len(worst))
            want = orc.apply_gate(host.copy(), qs, u)
            errs = []
            for tcv in (0, 1, 2, 3, 4, 5, 6):
                sim.set_tuning("tc", tcv)
                st = ss.Create(n); ss.from_numpy(host, st)
                sim.ApplyGate(qs, u, st)
                errs.append(float(np.abs(ss.to_numpy(st) - want).max()))
            print(json.dumps({"n": n, "G": g, "qs": qs, "err_cuda_cores": errs[0], "err_tc_smem": errs[1], "err_tc_smem_alt": errs[2], "err_tca": errs[3], "err_tca_nocomp": errs[4], "err_tca_ring3": errs[5], "err_tca_ring4": errs[6]}), flush=True)
            worst[g] = max(worst.get(g, 0), *errs[1:])
    # controlled
    sim.set_tuning("tc", 1)
    for qs, cqs, cv in (([3, 5, 6, 9], [1, 10], 0b10), ([0, 2, 4, 7], [11], 1), ([1, 2, 3, 4], [0], 1)):
        u = unitary(4, 99)
        st = ss.Create(n); ss.from_numpy(host, st)
        sim.ApplyControlledGate(qs, cqs, cv, u, st)
        err = float(np.abs(ss.to_numpy(st) - orc.apply_controlled_gate(host.copy(), qs, cqs, cv, u)).max())
        print(json.dumps({"n": n, "controlled": [qs, cqs, cv], "err_tc": err}), flush=True)
        worst["c"] = max(worst.get("c", 0), err)
print(json.dumps({"worst_err_tc": {str(k): v for k, v in worst.items()}}), flush=True)

if not args.skip_timing:
    n = args.n
    st = ss.Create(n); ss.SetStateUniform(st)
    for g in (4, 5):
        for qs in ([n - g + i for i in range(g)], [8 + i for i in range(g)], [5 + 4 * i for i in range(g)],
                   [3, 10, 13, 16, 19][:g], [0, 3, 7, 12, 29][:g], [0, 1, 2, 3, 4][:g], [1, 10, 13, 16, 19][:g]):
            u = unitary(g, 1)
            row = {"n": n, "G": g, "qs": qs}
            for tcv in (0, 1, 2, 3, 4, 5, 6):
                sim.set_tuning("tc", tcv)
                for _ in range(2): sim.ApplyGate(qs, u, st)
                ts = []
                for _ in range(7):
                    sim.timer_start(); sim.ApplyGate(qs, u, st); ts.append(sim.timer_stop_ms())
                row[["ms_cuda_cores", "ms_tc_smem", "ms_tc_smem_alt", "ms_tca", "ms_tca_nocomp", "ms_tca_ring3", "ms_tca_ring4"][tcv]] = round(float(np.median(ts)), 3)
            row["GBps_best_tc"] = round(16.0 * (1 << n) / min(row["ms_tca"], row["ms_tca_ring3"], row["ms_tca_ring4"]) / 1e6)
            print(json.dumps(row), flush=True)
    # norm drift over many passes (the tensor core's fp32 accumulation truncates): 64 random gates
    for g in (4, 5):
        for tcv in (0, 4, 3):
            sim.set_tuning("tc", tcv)
            ss.SetStateUniform(st)
            r2 = np.random.RandomState(3)
            for i in range(64):
                qs = sorted(r2.choice(np.arange(3, n), g, replace=False).tolist())
                sim.ApplyGate(qs, unitary(g, i), st)
            nrm = ss.Norm(st)
            print(json.dumps({"G": g, "tc": tcv, "norm_after_64_gates": nrm, "drift_per_pass": (nrm - 1) / 64}), flush=True)
