#!/usr/bin/env python
"""Per-kernel micro-benchmarks: one fused-gate pass per layout, CUDA-event
timed, median of `reps` after warm-up.  Prints one JSON line per case.

  python tools/microbench.py --n 30 [--dtype f32] [--tune gate_mode=0] [--out gpurun_out/mb.jsonl]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qsim_b200  # noqa: E402


def unitary(g, seed, cdt):
    rng = np.random.RandomState(seed)
    a = rng.standard_normal((1 << g, 1 << g)) + 1j * rng.standard_normal((1 << g, 1 << g))
    q, _ = np.linalg.qr(a)
    return q.astype(cdt)


def layouts(n, g):
    if g == 0:
        return {"phase": []}
    out = {"top": list(range(n - g, n)), "low": list(range(g)), "mid": list(range(8, 8 + g))}
    out["scattered"] = [5 + 4 * j for j in range(g)]
    out["mixed"] = sorted(set([0, 3, 7, 12, 17, n - 1][:g - 1] + [n - 1]))[:g] if g > 1 else [3]
    for q0 in (1, 2, 3, 4, 5, 6):
        out[f"q0={q0}"] = [q0] + [10 + 3 * j for j in range(g - 1)]
    return {k: v for k, v in out.items() if len(v) == g and max(v) < n}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=30)
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--gs", default="0,1,2,3,4,5,6")
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--tune", action="append", default=[])
    ap.add_argument("--out", default=None)
    ap.add_argument("--statespace", action="store_true")
    ap.add_argument("--controlled", action="store_true", help="controlled-gate sweep (C = 1..3, low/high/mixed controls)")
    args = ap.parse_args()
    rdt, cdt = (np.float32, np.complex64) if args.dtype == "f32" else (np.float64, np.complex128)
    ss, sim = qsim_b200.StateSpaceB200(rdt), qsim_b200.SimulatorB200(rdt)
    for t in args.tune:
        k, v = t.split("=")
        sim.set_tuning(k, int(v))
    n = args.n
    st = ss.Create(n)
    ss.SetStateUniform(st)
    amp_bytes = 8 if args.dtype == "f32" else 16
    pass_bytes = 2.0 * amp_bytes * (1 << n)
    fout = open(args.out, "a") if args.out else None

    def emit(rec):
        s = json.dumps(rec)
        print(s, flush=True)
        if fout:
            fout.write(s + "\n"); fout.flush()

    def timeit(fn, reps):
        for _ in range(2):
            fn()
        ts = []
        for _ in range(reps):
            sim.timer_start(); fn(); ts.append(sim.timer_stop_ms())
        return float(np.median(ts)), float(np.min(ts))

    for g in [int(x) for x in args.gs.split(",")]:
        for name, qs in layouts(n, g).items():
            u = unitary(g, g, cdt)
            med, best = timeit(lambda: sim.ApplyGate(qs, u, st), args.reps)
            emit({"op": "gate", "n": n, "dtype": args.dtype, "G": g, "layout": name, "qs": qs, "ms": med, "ms_min": best,
                  "GBps": pass_bytes / med / 1e6, "kernel": sim.last_kernel_name(), "tune": args.tune})
    # controlled gates: C = 1, 2, 3 controls on low / high / mixed bits, control values all ones and all
    # zeros, under targets of 1, 2 and 4 qubits (low and high).  Algorithmic bytes = 16 * 2^(n - C).
    ctl_sets = {1: {"low": [1], "high": [n - 2], "mixed": [6]},
                2: {"low": [1, 2], "high": [n - 3, n - 2], "mixed": [2, n - 2]},
                3: {"low": [1, 2, 4], "high": [n - 4, n - 3, n - 2], "mixed": [1, 11, n - 2]}}
    tgt_sets = {1: {"low": [0], "high": [n - 1], "mid": [9]},
                2: {"low": [0, 3], "high": [n - 5, n - 1], "mid": [8, 13]},
                4: {"low": [0, 3, 5, 7], "high": [n - 9, n - 7, n - 5, n - 1], "mid": [8, 9, 14, 15]}}
    if args.controlled:
        for g, tsets in tgt_sets.items():
            u = unitary(g, g, cdt)
            for tname, qs in tsets.items():
                for c, csets in ctl_sets.items():
                    for cname, cqs in csets.items():
                        if set(qs) & set(cqs) or max(qs + cqs) >= n:
                            continue
                        for cv_name, cv in (("ones", (1 << c) - 1), ("zeros", 0)):
                            med, best = timeit(lambda: sim.ApplyControlledGate(qs, cqs, cv, u, st), args.reps)
                            emit({"op": "cgate", "n": n, "G": g, "targets": tname, "C": c, "controls": cname, "cvals": cv_name,
                                  "qs": qs, "cqs": cqs, "ms": med, "kernel": sim.last_kernel_name(),
                                  "GBps": pass_bytes / (1 << c) / med / 1e6})
    sim2 = qsim_b200.SimulatorB200(rdt)
    for g, qs in ((1, [7]), (2, [3, 19]), (4, [8, 9, 14, 15]), (5, [0, 3, 7, 12, 20]), (6, [1, 5, 9, 13, 17, 21])):
        if max(qs) >= n:
            continue
        u = unitary(g, g, cdt)
        med, best = timeit(lambda: sim2.ExpectationValue(qs, u, st), max(3, args.reps // 2))
        emit({"op": "expect", "n": n, "G": g, "qs": qs, "ms": med, "GBps": pass_bytes / 2 / med / 1e6})
    if args.statespace:
        s2 = ss.Create(n)
        if not ss.IsNull(s2):
            ss.SetStateUniform(s2)
            for name, fn, nbytes in (("norm", lambda: ss.Norm(st), pass_bytes / 2),
                                     ("inner_product", lambda: ss.InnerProduct(st, s2), pass_bytes),
                                     ("multiply", lambda: ss.Multiply(1.0, st), pass_bytes),
                                     ("add", lambda: ss.Add(s2, st), pass_bytes * 1.5),
                                     ("set_uniform", lambda: ss.SetStateUniform(st), pass_bytes / 2),
                                     ("set_zero", lambda: ss.SetStateZero(st), pass_bytes / 2)):
                med, best = timeit(fn, 5)
                emit({"op": name, "n": n, "ms": med, "GBps": nbytes / med / 1e6})
            ss.SetStateUniform(st)
            for num in (1000, 100000):
                med, best = timeit(lambda: ss.Sample(st, num, 1), 3)
                emit({"op": "sample", "n": n, "num_samples": num, "ms": med})
            med, best = timeit(lambda: ss.Measure([0, n // 2, n - 1], 0.3, st), 3)
            emit({"op": "measure", "n": n, "ms": med})


if __name__ == "__main__":
    main()
