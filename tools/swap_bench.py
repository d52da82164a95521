#!/usr/bin/env python
"""Micro-benchmark of the in-place local<->global swap kernel (csrc/p2p_swap.cu) under torchrun:
k swapped local bits at chosen positions, GB/s per direction per GPU.
  torchrun --nproc-per-node N tools/swap_bench.py [n_local]"""
import json, os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qsim_b200.sharded import B200Engine

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_local = int(sys.argv[1]) if len(sys.argv) > 1 else 30
g = world.bit_length() - 1
eng = B200Engine(n_local, local, p2p=True)
eng.connect_peers(dist, rank, world)
eng.ss.SetStateUniform(eng.state)
cases = []
for k in range(1, g + 1):
    tops = list(range(n_local - k, n_local))
    cases += [(k, tops), (k, list(range(12, 12 + k))), (k, list(range(5, 5 + k))), (k, list(range(2, 2 + k))),
              (k, [3 + 9 * j for j in range(k)])]
for k, lbits in cases:
    gbits = list(range(k))  # swap with the k lowest rank bits
    my = sum(((rank >> gb) & 1) << j for j, gb in enumerate(gbits))
    def peer(b):
        r = rank
        for j, gb in enumerate(gbits):
            r = (r & ~(1 << gb)) | (((b >> j) & 1) << gb)
        return r
    peers = [None if b == my else peer(b) for b in range(1 << k)]
    ts = []
    for it in range(4):
        eng.stream_barrier(dist)
        e0 = eng.event()
        eng.swap_global_local(peers, k, lbits, my)
        e1 = eng.event()
        eng.stream_barrier(dist)
        torch.cuda.synchronize()
        if it:
            ts.append(e0.elapsed_time(e1))
    t = torch.tensor([float(np.median(ts))], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sent = (8 << n_local) * ((1 << k) - 1) / (1 << k)
    if rank == 0:
        print(json.dumps({"world": world, "n_local": n_local, "k": k, "local_bits": lbits, "ms": round(float(t.item()), 3),
                          "GBps_per_direction": round(sent / float(t.item()) / 1e6, 1)}), flush=True)
dist.barrier()
dist.destroy_process_group()
