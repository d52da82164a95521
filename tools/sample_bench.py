#!/usr/bin/env python
"""Sample(state, m, seed) with the random values drawn on the device (csrc/sample_rng.cu) against the reference's
flow (host mt19937 + std::sort + copy): python tools/sample_bench.py [n] -> one JSON line per sample count."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qsim_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
ss = qsim_b200.StateSpaceB200(np.float32)
st = ss.Create(n)
ss.SetStateUniform(st)
for num in (1000, 100000, 1000000):
    res = {}
    for name, host in (("device_rng", False), ("host_rng", True)):
        ss.Sample(st, num, 1, host_rng=host)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            out = ss.Sample(st, num, 1, host_rng=host)
            ts.append((time.perf_counter() - t0) * 1e3)
        res[name + "_ms"] = round(min(ts), 3)
        res[name + "_first"] = int(out[0])
        res.setdefault("outs", []).append(out)
    same = bool(np.array_equal(*res.pop("outs")))
    t0 = time.perf_counter()
    ss.Norm(st)
    res["norm_ms"] = round((time.perf_counter() - t0) * 1e3, 3)
    print(json.dumps({"n": n, "samples": num, "identical_indices": same, **res}), flush=True)
