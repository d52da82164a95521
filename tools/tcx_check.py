#!/usr/bin/env python
"""Checks k_gate_tcx (tensor-core 6-qubit gates and 4/5/6-qubit expectation values) against the oracle at small n,
times it against the FFMA2 kernels at n=30 and measures the norm drift of the G=6 gate for the compensation constant."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qsim_b200
from oracle.oracle import Oracle

orc = Oracle()
ss, sim = qsim_b200.StateSpaceB200(np.float32), qsim_b200.SimulatorB200(np.float32)
rng = np.random.RandomState(5)

def unitary(g, seed):
    r = np.random.RandomState(seed)
    a = r.standard_normal((1 << g, 1 << g)) + 1j * r.standard_normal((1 << g, 1 << g))
    u, _ = np.linalg.qr(a)
    return u.astype(np.complex64)

worst = {}
for n in (13, 15):
    host = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    host = (host / np.linalg.norm(host)).astype(np.complex64)
    for g in (4, 5, 6):
        layouts = [list(range(g)), list(range(n - g, n)), list(range(3, 3 + g)), [0] + list(range(n - g + 1, n)),
                   sorted(rng.choice(n, g, replace=False).tolist()), sorted(rng.choice(n, g, replace=False).tolist())]
        for qs in layouts:
            u = unitary(g, g + len(qs) + qs[0])
            st = ss.Create(n); ss.from_numpy(host, st)
            want = orc.expectation_value(host, qs, u)
            sim.set_tuning("tcx", 0); e0 = sim.ExpectationValue(qs, u, st)
            sim.set_tuning("tcx", -1); e1 = sim.ExpectationValue(qs, u, st)
            worst[f"expect{g}"] = max(worst.get(f"expect{g}", 0), abs(e1 - want))
            rec = {"n": n, "G": g, "qs": qs, "expect_err_ffma": abs(e0 - want), "expect_err_tc": abs(e1 - want)}
            if g == 6:
                wantg = orc.apply_gate(host.copy(), qs, u)
                sim.ApplyGate(qs, u, st)
                rec["gate_err_tc"] = float(np.abs(ss.to_numpy(st) - wantg).max())
                worst["gate6"] = max(worst.get("gate6", 0), rec["gate_err_tc"])
            print(json.dumps(rec), flush=True)
print(json.dumps({"worst": worst}), flush=True)

n = 30
st = ss.Create(n); ss.SetStateUniform(st)
for g in (4, 5, 6):
    for qs in ([n - g + i for i in range(g)], [5 + 4 * i for i in range(g)], [0, 3, 7, 12, 17, 29][:g], [1, 10, 13, 16, 19, 22][:g]):
        u = unitary(g, 1)
        row = {"n": n, "G": g, "qs": qs}
        for name, v in (("ffma", 0), ("tc", -1)):
            sim.set_tuning("tcx", v)
            ts = []
            for _ in range(4):
                sim.timer_start(); sim.ExpectationValue(qs, u, st); ts.append(sim.timer_stop_ms())
            row["expect_ms_" + name] = round(float(np.median(ts[1:])), 3)
            if g == 6:
                ts = []
                for _ in range(4):
                    sim.timer_start(); sim.ApplyGate(qs, u, st); ts.append(sim.timer_stop_ms())
                row["gate_ms_" + name] = round(float(np.median(ts[1:])), 3)
        print(json.dumps(row), flush=True)
for comp in (0, 300):
    sim.set_tuning("tcx", -1); sim.set_tuning("tc_comp6", comp)
    ss.SetStateUniform(st)
    r2 = np.random.RandomState(3)
    for i in range(32):
        qs = sorted(r2.choice(np.arange(3, n), 6, replace=False).tolist())
        sim.ApplyGate(qs, unitary(6, i), st)
    nrm = ss.Norm(st)
    print(json.dumps({"G": 6, "tc_comp6": comp, "norm_after_32_gates": nrm, "drift_per_pass": (nrm - 1) / 32}), flush=True)
