#!/usr/bin/env python
"""Micro-benchmark of the local<->global exchange of the sharded state (csrc/sharded.cu: k_remap_push out of
place, k_p2p_swap in place): k victims at chosen local bits, GB/s per direction per GPU, device-timed by the
library (CUDA events around kernel + barrier on the first local shard).

  python tools/swap_bench.py --single N [n_local]        one process drives N GPUs (event barriers)
  torchrun --nproc-per-node N tools/swap_bench.py [n_local]   one process per GPU (flag barriers)"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qsim_b200.sv import ShardedStateB200  # noqa: E402

args = sys.argv[1:]
single = 0
if args and args[0] == "--single":
    single = int(args[1])
    args = args[2:]
n_local = int(args[0]) if args else 30
dist = None
if single:
    rank, world = 0, single
    g = world.bit_length() - 1
    sv = ShardedStateB200.single_process(list(range(single)), n_local + g)
else:
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g = world.bit_length() - 1
    sv = ShardedStateB200.multi_process(dist, n_local + g, local)
n = n_local + g
for kv in os.environ.get("QB200_SV_OPTIONS", "").split(","):   # e.g. push_kernel=1,push_ctas_per_sm=2
    if kv:
        sv.set_option(kv.split("=")[0], int(kv.split("=")[1]))
sv.SetStateUniform()
cases = []
for k in range(1, g + 1):
    cases += [(k, list(range(n_local - k, n_local))), (k, list(range(12, 12 + k))), (k, list(range(5, 5 + k))),
              (k, list(range(2, 2 + k))), (k, list(range(0, k))), (k, [3 + 9 * j for j in range(k)])]
for mode in (1, 0):
    sv.set_option("swap_mode", mode)
    for k, lbits in cases:
        ts = []
        for it in range(4):
            pos = sv.qubit_map()
            at = {p: q for q, p in enumerate(pos)}
            victims = [at[b] for b in lbits]
            incoming = [at[n_local + t] for t in range(k)]
            sv.reset_stats()
            sv.Swap(victims, incoming)
            st = sv.stats()
            if it:
                ts.append(st["exchange_ms"] / max(1, st["swaps"]))
        t = float(np.median(ts))
        if dist is not None:
            tt = torch.tensor([t], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = float(tt.item())
        sent = (8 << n_local) * ((1 << k) - 1) / (1 << k)
        if rank == 0:
            print(json.dumps({"world": world, "mode": "out-of-place push" if mode else "in-place pull+push", "n_local": n_local, "k": k,
                              "victim_bits": lbits, "local_swap_passes": st["local_swap_passes"], "ms": round(t, 3),
                              "GBps_per_direction": round(sent / t / 1e6, 1)}), flush=True)
sv.close()
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
