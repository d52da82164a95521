#!/usr/bin/env python
"""A few passes through every fp32/fp64 gate kernel at small n, for `compute-sanitizer --tool memcheck|racecheck`."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qsim_b200

def unitary(g, seed, cdt):
    r = np.random.RandomState(seed)
    a = r.standard_normal((1 << g, 1 << g)) + 1j * r.standard_normal((1 << g, 1 << g))
    u, _ = np.linalg.qr(a)
    return u.astype(cdt)

n = 15
for rdt, cdt in ((np.float32, np.complex64), (np.float64, np.complex128)):
    ss, sim = qsim_b200.StateSpaceB200(rdt), qsim_b200.SimulatorB200(rdt)
    st = ss.Create(n); ss.SetStateUniform(st)
    for qs in ([3], [0, 7], [1, 4, 9], [5, 8, 11, 14], [0, 6, 9, 12], [0, 1, 2, 3], [1, 2, 7, 8], [4, 6, 8, 10, 12],
               [0, 3, 5, 9, 13], [2, 4, 6, 8, 10, 12], [0, 1, 5, 9, 11, 14]):
        u = unitary(len(qs), len(qs), cdt)
        sim.ApplyGate(qs, u, st)
        sim.ExpectationValue(qs, u, st)
    sim.ApplyControlledGate([5, 9], [2, 12], 0b01, unitary(2, 7, cdt), st)
    print(rdt.__name__, "norm", ss.Norm(st), "samples", ss.Sample(st, 8, 1)[:3], flush=True)
    ss.Measure([0, 7], 0.3, st)
    # batched reductions (mapped pinned result slots) and the all-qubit moments kernel (3 passes at n = 15 / 14)
    terms = [(qs, unitary(len(qs), 3, cdt)) for qs in ([2], [0, 9], [1, 5, 9, 13, 14], [0, 1, 5, 9, 11, 14])] * 20
    vals = sim.ExpectationValues(terms, st)
    # XOR-monomial operators (expect_monomial.cu): diagonal, in-pair swap, general; bit 0 in and out of the mask
    P = {"X": np.array([[0, 1], [1, 0]]), "Y": np.array([[0, -1j], [1j, 0]]), "Z": np.diag([1, -1])}
    for names, qs in (("ZZZ", [0, 5, 14]), ("XZZ", [0, 5, 14]), ("XZYXZY", [0, 1, 2, 3, 4, 5]), ("ZXY", [3, 9, 14]),
                      ("YYYYY", [1, 2, 6, 10, 13])):
        m = np.array([[1.0]])
        for c in names:
            m = np.kron(P[c], m)
        sim.ExpectationValue(qs, m.astype(cdt), st)
    mom = sim.OneQubitMoments(st)
    st13 = ss.Create(13); ss.SetStateUniform(st13)
    mom13 = sim.OneQubitMoments(st13)
    st3 = ss.Create(3); ss.SetStateUniform(st3)
    mom3 = sim.OneQubitMoments(st3)
    print(rdt.__name__, "batched", len(vals), "moments", mom.shape, float(mom[:, :2].sum(axis=1).max()), mom13.shape, mom3.shape, flush=True)
    # round 2: device-side random values + seeded sampler, several operators in one pass
    rs_dev = ss.GenerateRandomValuesOnDevice(1000, 5, 0.9)
    smp = ss.Sample(st, 700, 3)
    multi = sim.ExpectationValuesSameQubits([2, 9], [unitary(2, k, cdt) for k in range(5)], st)
    multi1 = sim.ExpectationValuesSameQubits([0], [unitary(1, k, cdt) for k in range(8)], st)
    print(rdt.__name__, "device rng", float(rs_dev[0]), int(smp[0]), "multi", multi.shape, multi1.shape, flush=True)
ss.DeviceSync()

# sharded state on one device (4 shards): both push kernels, the in-place kernel, the copy-engine path and the
# overlapped pipeline (chunked controlled passes + slim push on a second stream)
from qsim_b200.sv import ShardedStateB200
from qsim_b200.trace import TraceOp
rs = np.random.RandomState(3)
for rdt, cdt in ((np.float32, np.complex64), (np.float64, np.complex128)):
    n = 16
    ops = []
    for i in range(24):
        g = int(rs.randint(1, 5))
        qs = sorted(rs.choice(n, g, replace=False).tolist())
        ops.append(TraceOp(qs, [], 0, np.ascontiguousarray(unitary(g, i, np.complex64)).reshape(-1).view(np.float32).copy()))
    for opts in ({"push_kernel": 0}, {"push_kernel": 1}, {"push_kernel": 2}, {"swap_mode": 0},
                 {"overlap": 1, "overlap_ce": 0}, {"overlap": 1, "overlap_ce": 1, "overlap_chunks_log2": 1}):
        sv = ShardedStateB200.single_process([0] * 4, n, rdt)
        for k, v in opts.items():
            sv.set_option(k, v)
        sv.SetStateZero()
        sv.Run(ops)
        pos = sv.qubit_map()
        at = {p: q for q, p in enumerate(pos)}
        sv.Swap([at[12], at[13]], [at[14], at[15]])   # victims at bits 12, 13: the copy-engine path when asked for
        print(rdt.__name__, opts, "norm", round(sv.Norm(), 6), sv.stats()["swaps"], sv.stats()["copy_engine_swaps"], flush=True)
        sv.close()
print("done")
