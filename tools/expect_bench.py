#!/usr/bin/env python
"""Read-only passes: ExpectationValue per layout and Norm, per call (one stream
synchronisation each, wall clock over `reps` calls) and batched (SimulatorB200.ExpectationValues:
all passes enqueued, one synchronisation).  One JSON line per case.

  python tools/expect_bench.py --n 30 [--reps 20]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qsim_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=30)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--tune", action="append", default=[], help="key=value for qb200_ctx_set_tuning")
ap.add_argument("--small", action="store_true", help="G <= 2 layouts only")
args = ap.parse_args()
n = args.n
ss, sim = qsim_b200.StateSpaceB200(np.float32), qsim_b200.SimulatorB200(np.float32)
for t in args.tune:
    k, v = t.split("=")
    sim.set_tuning(k, int(v))
st = ss.Create(n)
ss.SetStateUniform(st)
rng = np.random.RandomState(3)
read_bytes = 8.0 * (1 << n)


def matrix(g):
    return (rng.standard_normal((1 << g, 1 << g)) + 1j * rng.standard_normal((1 << g, 1 << g))).astype(np.complex64)


cases = [[7], [0], [1], [n - 1], [3, 19], [0, 1], [0, 9], [2, 5, 11], [8, 9, 14, 15], [0, 3, 7, 12, 20],
         [1, 5, 9, 13, 17, 21]]
if args.small:
    cases = [qs for qs in cases if len(qs) <= 2]
for qs in cases:
    m = matrix(len(qs))
    for _ in range(3):
        sim.ExpectationValue(qs, m, st)
    t0 = time.perf_counter()
    for _ in range(args.reps):
        v = sim.ExpectationValue(qs, m, st)
    single = (time.perf_counter() - t0) / args.reps * 1e3
    terms = [(qs, m)] * args.reps
    sim.ExpectationValues(terms[:3], st)
    t0 = time.perf_counter()
    vb = sim.ExpectationValues(terms, st)
    batch = (time.perf_counter() - t0) / args.reps * 1e3
    assert vb[-1] == v
    print(json.dumps({"n": n, "tune": args.tune, "expect": qs, "ms_per_call": round(single, 4), "GBps": round(read_bytes / single / 1e6),
                      "ms_batched": round(batch, 4), "GBps_batched": round(read_bytes / batch / 1e6)}), flush=True)
# Pauli strings (XOR-monomial matrices): the read pass of csrc/expect_monomial.cu against the dense kernels (mono=0)
PAULI = {"X": np.array([[0, 1], [1, 0]]), "Y": np.array([[0, -1j], [1j, 0]]), "Z": np.diag([1, -1])}
dense = qsim_b200.SimulatorB200(np.float32)
dense.set_tuning("mono", 0)
if not args.small:
    for names, qs in (("XZY", [2, 5, 11]), ("XZYX", [8, 9, 14, 15]), ("ZZZZZ", [0, 3, 7, 12, 20]),
                      ("XZYXZY", [0, 1, 2, 3, 4, 5]), ("XZYXZY", [1, 5, 9, 13, 17, 21]), ("XZYXZY", [n - 6 + k for k in range(6)])):
        m = np.array([[1.0]])
        for c in names:
            m = np.kron(PAULI[c], m)
        m = m.astype(np.complex64)
        rec = {"n": n, "pauli": names, "qs": qs}
        for label, s_ in (("read_pass", sim), ("dense", dense)):
            terms = [(qs, m)] * args.reps
            s_.ExpectationValues(terms[:3], st)
            t0 = time.perf_counter()
            s_.ExpectationValues(terms, st)
            ms = (time.perf_counter() - t0) / args.reps * 1e3
            rec[f"ms_batched_{label}"] = round(ms, 4)
            rec[f"GBps_{label}"] = round(read_bytes / ms / 1e6)
        print(json.dumps(rec), flush=True)
for _ in range(3):
    ss.Norm(st)
t0 = time.perf_counter()
for _ in range(args.reps):
    ss.Norm(st)
ms = (time.perf_counter() - t0) / args.reps * 1e3
print(json.dumps({"n": n, "norm_ms": round(ms, 4), "GBps": round(read_bytes / ms / 1e6)}), flush=True)
