#!/usr/bin/env python
"""Feasibility probe for L2 cache blocking (chunk-major execution): a sequence of fused gates whose targets all
sit below bit c is applied (a) gate by gate over the whole 2^n state (every pass streams HBM) and (b) chunk by
chunk -- all gates on the 2^c-amplitude chunk 0, then chunk 1, ... -- so that a chunk stays in the 126 MB L2
between its gates.  Both orders are captured into CUDA graphs (launch overhead of the host is not what is
measured) and timed with CUDA events.  Same kernels, same results.

  python tools/l2_probe.py [--n 28] [--gates 8]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qsim_b200  # noqa: E402


def unitary(g, seed):
    rng = np.random.RandomState(seed)
    a = rng.standard_normal((1 << g, 1 << g)) + 1j * rng.standard_normal((1 << g, 1 << g))
    q, _ = np.linalg.qr(a)
    return q.astype(np.complex64)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=28)
    ap.add_argument("--gates", type=int, default=8)
    ap.add_argument("--chunks", default="20,21,22,23,24")
    args = ap.parse_args()
    n = args.n
    ss, sim = qsim_b200.StateSpaceB200(np.float32), qsim_b200.SimulatorB200(np.float32)
    stream = torch.cuda.Stream()
    sim.set_stream(stream.cuda_stream)
    ss.set_stream(stream.cuda_stream)
    st = ss.Create(n)
    ss.SetStateUniform(st)
    rng = np.random.RandomState(1)
    out = []
    for c in [int(x) for x in args.chunks.split(",")]:
        # gates inside the low c bits: G = 4 with a lowest target >= 4 (tensor-core kernel) and a few G = 2
        gates = []
        for i in range(args.gates):
            g = 4 if i % 4 != 3 else 2
            qs = sorted(rng.choice(np.arange(4, c), g, replace=False).tolist())
            gates.append((qs, unitary(g, 10 * c + i)))
        chunks = [ss.CreateFromPointer(st.get() + (8 << c) * k, c) for k in range(1 << (n - c))]

        def gate_major():
            for qs, u in gates:
                sim.ApplyGate(qs, u, st)

        def chunk_major():
            for ch in chunks:
                for qs, u in gates:
                    sim.ApplyGate(qs, u, ch)

        res = {"n": n, "chunk_qubits": c, "chunk_MiB": (8 << c) / 2**20, "gates": len(gates)}
        for name, fn in (("gate_major", gate_major), ("chunk_major", chunk_major)):
            with torch.cuda.stream(stream):
                fn()   # warm-up: function attributes, lazy kernel loading
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=stream):
                    fn()
                ts = []
                for _ in range(4):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    graph.replay()
                    e1.record(stream)
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
            ms = float(np.median(ts[1:]))
            res[name + "_ms"] = ms
            res[name + "_ms_per_gate_pass"] = ms / len(gates)
            res[name + "_algorithmic_GBps"] = len(gates) * 16.0 * (1 << n) / ms / 1e6
        res["speedup"] = res["gate_major_ms"] / res["chunk_major_ms"]
        print(json.dumps(res), flush=True)
        out.append(res)


if __name__ == "__main__":
    main()
