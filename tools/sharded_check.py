#!/usr/bin/env python
"""Multi-GPU parity check of the sharded state of the C ABI (qb200_sv_*).

  torchrun --nproc-per-node N tools/sharded_check.py [trace]      one process per GPU (multi-process mode)
  python tools/sharded_check.py --single N [trace]                 one process drives N GPUs

The fused trace (default: circuit_q24 depth 20) on a state sharded over N GPUs must equal the single-GPU result
(rank 0 recomputes it unsharded): 64 amplitudes, norm, expectation values on local and global qubits, sampling
with the same seed.  Prints one JSON line from rank 0; exit code 1 on mismatch."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qsim_b200  # noqa: E402
from qsim_b200.sv import ShardedStateB200  # noqa: E402

args = sys.argv[1:]
single = 0
if args and args[0] == "--single":
    single = int(args[1])
    args = args[2:]
trace = args[0] if args else "tests/golden/q24_d20_f4.trace"
n, ops = qsim_b200.read_trace(trace)

dist = None
if single:
    rank, world, local = 0, single, 0
    sv = ShardedStateB200.single_process(list(range(single)), n)
else:
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sv = ShardedStateB200.multi_process(dist, n, local)
for kv in os.environ.get("QB200_SV_OPTIONS", "").split(","):
    if kv:
        sv.set_option(kv.split("=")[0], int(kv.split("=")[1]))

sv.SetStateZero()
sv.Run(ops)
norm = sv.Norm()
rs = np.random.RandomState(0)
idx = [0, 1, 2, 7] + rs.randint(0, 1 << n, 60).tolist()
amps = [sv.GetAmpl(i) for i in idx]
P = {"X": np.array([[0, 1], [1, 0]]), "Y": np.array([[0, -1j], [1j, 0]]), "Z": np.diag([1, -1])}


def pauli(names):
    m = np.array([[1.0]])
    for c in names:
        m = np.kron(P[c], m)
    return m.astype(np.complex64)


ecases = [([0], pauli("X")), ([n - 1], pauli("Z")), ([1, n - 1], rs.standard_normal((4, 4)).astype(np.complex64)),
          ([2, 9, n - 2, n - 1], pauli("XZYX")), ([0, 3, 7, 12, n - 3, n - 1], pauli("XZYXZY"))]
st0 = sv.stats()
evs = [sv.ExpectationValue(qs, m) for qs, m in ecases]
st1 = sv.stats()
samples = sv.Sample(256, 11)
pos_final = sv.qubit_map()
ok, maxerr, everr, same_samples = True, 0.0, 0.0, None
if rank == 0:
    ss, s1 = qsim_b200.StateSpaceB200(np.float32, device=local), qsim_b200.SimulatorB200(np.float32, device=local)
    st = ss.Create(n)
    ss.SetStateZero(st)
    for op in ops:
        if op.controls:
            s1.ApplyControlledGate(op.qubits, op.controls, op.cvals, op.matrix, st)
        else:
            s1.ApplyGate(op.qubits, op.matrix, st)
    for i, a in zip(idx, amps):
        maxerr = max(maxerr, abs(a - ss.GetAmpl(st, i)))
    for (qs, m), v in zip(ecases, evs):
        everr = max(everr, abs(v - s1.ExpectationValue(qs, m, st)))
    same_samples = float(np.mean(ss.Sample(st, 256, 11) == samples))
    ok = maxerr < 1e-6 and abs(norm - 1) < 1e-4 and everr < 1e-5 and same_samples > 0.95 and st0["swaps"] >= 1
    print(json.dumps({"world": world, "mode": "single-process" if single else "multi-process", "n": n, "ops": len(ops),
                      "swaps": st0["swaps"], "swaps_for_expectations": st1["swaps"] - st0["swaps"],
                      "max_abs_err_expectations": everr, "local_swap_passes": st1["local_swap_passes"],
                      "bytes_sent_per_rank": st0["bytes_sent_per_shard"], "norm": norm, "max_abs_err_vs_single_gpu": maxerr,
                      "same_samples_fraction": same_samples, "ok": bool(ok), "exchange_ms": st0["exchange_ms"],
                      "final_qubit_map": pos_final}))
sv.close()
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
