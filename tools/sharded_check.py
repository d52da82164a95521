#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun on N GPUs): the q24 depth-20 fused trace on a
state sharded over N ranks must equal the single-GPU result (rank 0 recomputes it unsharded).
Prints one JSON line from rank 0."""
import json, os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qsim_b200
from qsim_b200.sharded import B200Engine, ShardedSimulator

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
trace = sys.argv[1] if len(sys.argv) > 1 else "tests/golden/q24_d20_f4.trace"
n, ops = qsim_b200.read_trace(trace)
g = world.bit_length() - 1
p2p = os.environ.get("QB200_P2P", "1") == "1"
eng = B200Engine(n - g, local, p2p=p2p)
if p2p:
    eng.connect_peers(dist, rank, world)
sim = ShardedSimulator(n, eng, dist=dist, rank=rank, world_size=world, transfer_scalars=1 << 20)
sim.set_state_zero()
plan = sim.run(ops)
norm = sim.norm()
rs = np.random.RandomState(0)
idx = [0, 1, 2, 7] + rs.randint(0, 1 << n, 60).tolist()
amps = [sim.get_ampl(i) for i in idx]
# expectation values on the sharded state: Pauli strings (read pass) and dense operators, local and global qubits
P = {"X": np.array([[0, 1], [1, 0]]), "Y": np.array([[0, -1j], [1j, 0]]), "Z": np.diag([1, -1])}
def pauli(names):
    m = np.array([[1.0]])
    for c in names:
        m = np.kron(P[c], m)
    return m.astype(np.complex64)
ecases = [([0], pauli("X")), ([n - 1], pauli("Z")), ([1, n - 1], rs.standard_normal((4, 4)).astype(np.complex64)),
          ([2, 9, n - 2, n - 1], pauli("XZYX")), ([0, 3, 7, 12, n - 3, n - 1], pauli("XZYXZY"))]
swaps_run = sim.stats.swaps
evs = [sim.expectation_value(qs, m) for qs, m in ecases]
ok, maxerr, everr = True, 0.0, 0.0
if rank == 0:
    ss, s1 = qsim_b200.StateSpaceB200(np.float32, device=local), qsim_b200.SimulatorB200(np.float32, device=local)
    st = ss.Create(n); ss.SetStateZero(st)
    for op in ops:
        s1.ApplyGate(op.qubits, op.matrix, st)
    for i, a in zip(idx, amps):
        maxerr = max(maxerr, abs(a - ss.GetAmpl(st, i)))
    for (qs, m), v in zip(ecases, evs):
        everr = max(everr, abs(v - s1.ExpectationValue(qs, m, st)))
    ok = maxerr < 1e-6 and abs(norm - 1) < 1e-4 and everr < 1e-5
    print(json.dumps({"world": world, "n": n, "ops": len(ops), "swaps": swaps_run, "swaps_for_expectations": sim.stats.swaps - swaps_run,
                      "max_abs_err_expectations": everr, "local_swap_passes": sim.stats.local_swap_passes,
                      "bytes_sent_per_rank": sim.stats.bytes_sent, "norm": norm, "max_abs_err_vs_single_gpu": maxerr, "ok": ok, "p2p": p2p,
                      "exchange_ms": sim.exchange_device_ms(), "final_global_qubits": sim.global_qubits()}))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
