#!/usr/bin/env python
"""Dispatch table sweep: for every choice of target bits inside the warp's lane range (bits
0..4; the remaining targets are spread over high bits) time a G = 4 pass on the three fp32 kernels that can take
it -- tensor cores (k_gate_tca, tuning tc=3), warp tile (tile=2, tc=0), cp.async ring (tile=1, tc=0) -- and a G = 5
pass on k_gate_tca<5> (tc=3) and the FFMA2 row-block kernel (tc=0).  Prints one JSON line per (n, G, low-bit mask)
with the winner; profiles/r02_dispatch_sweep.txt is the input of the table in csrc/gate_launch.cuh.
  python tools/dispatch_sweep.py [--n 30] [--reps 5]"""
import argparse
import itertools
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qsim_b200  # noqa: E402


def unitary(g, seed):
    rng = np.random.RandomState(seed)
    a = rng.standard_normal((1 << g, 1 << g)) + 1j * rng.standard_normal((1 << g, 1 << g))
    q, _ = np.linalg.qr(a)
    return q.astype(np.complex64)


ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=30)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--gs", default="4,5")
args = ap.parse_args()
n = args.n
ss = qsim_b200.StateSpaceB200(np.float32)
st = ss.Create(n)
ss.SetStateUniform(st)
VARIANTS = {4: {"tca": {"tc": 3}, "tile": {"tc": 0, "tile": 2}, "pipe": {"tc": 0, "tile": 1}},
            5: {"tca": {"tc": 3}, "big": {"tc": 0}}}
sims = {}
for g, vs in VARIANTS.items():
    for name, tune in vs.items():
        sim = qsim_b200.SimulatorB200(np.float32)
        for k, v in tune.items():
            sim.set_tuning(k, v)
        sims[(g, name)] = sim


def timeit(sim, qs, u):
    for _ in range(2):
        sim.ApplyGate(qs, u, st)
    ts = []
    for _ in range(args.reps):
        sim.timer_start()
        sim.ApplyGate(qs, u, st)
        ts.append(sim.timer_stop_ms())
    return float(np.median(ts))


for g in [int(x) for x in args.gs.split(",")]:
    u = unitary(g, g)
    for nlow in range(0, g + 1):
        for low in itertools.combinations(range(5), nlow):
            for spread in ("near", "far"):
                # remaining targets: "near" = directly above the lane bits (5, 6, ...), "far" = spread over high bits
                rest = [5 + j for j in range(g - nlow)] if spread == "near" else [10 + 3 * j for j in range(g - nlow)]
                qs = sorted(list(low) + rest)
                if max(qs) >= n:
                    continue
                res = {"n": n, "G": g, "low_mask": sum(1 << b for b in low), "low_bits": list(low), "spread": spread, "qs": qs}
                for name in VARIANTS[g]:
                    sim = sims[(g, name)]
                    res[name + "_ms"] = timeit(sim, qs, u)
                    res[name + "_kernel"] = sim.last_kernel_name()
                res["best"] = min(VARIANTS[g], key=lambda k: res[k + "_ms"])
                print(json.dumps(res), flush=True)
