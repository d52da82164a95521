#!/usr/bin/env python
"""Launches a few fused-gate passes (for ncu captures).  usage: one_gate.py n q0,q1,.. [f64] [key=val ...]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qsim_b200
n = int(sys.argv[1]); qs = [int(x) for x in sys.argv[2].split(",")]; g = len(qs)
f64 = "f64" in sys.argv[3:]
rdt, cdt = (np.float64, np.complex128) if f64 else (np.float32, np.complex64)
ss, sim = qsim_b200.StateSpaceB200(rdt), qsim_b200.SimulatorB200(rdt)
for t in [a for a in sys.argv[3:] if a != "f64"]:
    k, v = t.split("="); sim.set_tuning(k, int(v))
rng = np.random.RandomState(1)
a = rng.standard_normal((1 << g, 1 << g)) + 1j * rng.standard_normal((1 << g, 1 << g))
u, _ = np.linalg.qr(a); u = u.astype(cdt)
st = ss.Create(n); ss.SetStateUniform(st)
for _ in range(4): sim.ApplyGate(qs, u, st)
ss.DeviceSync()
