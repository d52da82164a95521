#!/usr/bin/env python
"""Samples SM clock / power while a gate kernel runs back to back (tools only)."""
import os, subprocess, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qsim_b200

def unitary(g, seed):
    rng = np.random.RandomState(seed)
    a = rng.standard_normal((1 << g, 1 << g)) + 1j * rng.standard_normal((1 << g, 1 << g))
    q, _ = np.linalg.qr(a)
    return q.astype(np.complex64)

n = 30
ss, sim = qsim_b200.StateSpaceB200(np.float32), qsim_b200.SimulatorB200(np.float32)
for t in sys.argv[1:]:
    k, v = t.split("="); sim.set_tuning(k, int(v))
st = ss.Create(n); ss.SetStateUniform(st)
for g, qs in ((2, [8, 9]), (4, [8, 9, 14, 15]), (5, [8, 9, 14, 15, 20])):
    u = unitary(g, g)
    for _ in range(3): sim.ApplyGate(qs, u, st)
    ss.DeviceSync()
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks.mem", "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
    time.sleep(0.15)
    reps = 250 if g < 5 else 100
    sim.timer_start()
    for _ in range(reps): sim.ApplyGate(qs, u, st)
    ms = sim.timer_stop_ms()
    p.terminate(); out = p.communicate()[0].strip().splitlines()
    rows = [r.split(",") for r in out]
    clk = [float(r[0]) for r in rows[5:]]; pw = [float(r[1]) for r in rows[5:]]
    print(f"G={g}: {ms/reps:.3f} ms/launch; sm clock median {np.median(clk):.0f} min {min(clk):.0f} MHz; power median {np.median(pw):.0f} max {max(pw):.0f} W; power_cap {[r[2].strip() for r in rows[5:]].count('Active')}/{len(rows)-5}; n_samples {len(clk)}")
