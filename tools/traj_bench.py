#!/usr/bin/env python
"""BASELINE config 5: 26-qubit depth-20 RQC (tools/gen_rqc.py, circuit_q30 rules) with depolarizing
noise p=0.001 after every gate qubit; trajectories split over N GPUs; observables X_q, Z_q on every
qubit + one 6-qubit Pauli string per window.  Prints one JSON line (trajectories/s, aggregate
algorithmic HBM GB/s).  A bounded sample of the 8192 repetition ids is run (--num per GPU).

  python tools/traj_bench.py --gpus 1 --num 64 [--n 26] [--depth 20] [--fused 4]
"""
import argparse
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qsim_b200.traj_farm import run_farm  # noqa: E402
from tools.gen_rqc import generate  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--num", type=int, default=64, help="trajectories per GPU")
ap.add_argument("--n", type=int, default=26)
ap.add_argument("--depth", type=int, default=20)
ap.add_argument("--fused", type=int, default=4)
ap.add_argument("--p", type=float, default=0.001)
ap.add_argument("--procs-per-gpu", type=int, default=1,
                help="worker processes per GPU (time-sliced: fills the host-side gaps between a trajectory's launches)")
ap.add_argument("--batch", type=int, default=2,
                help="2: single-qubit observables from the reduced density matrices (csrc/moments.cu) + the rest in one "
                     "batch; 1: all observables of a trajectory in one batch (expect_b200.h, one stream synchronisation); "
                     "0: the reference's lib/expect.h loop, one synchronisation per operator string")
ap.add_argument("--prefix", type=int, default=1,
                help="1: trajectories share the noiseless prefix of the fused gate list (include/qsim_b200/qtrajectory_b200.h)")
ap.add_argument("--total", type=int, default=0, help="total number of repetition ids over all GPUs (overrides --num), e.g. 8192")
ap.add_argument("--binary", default=None, help="worker binary (default apps/_bin/qsim_qtrajectory_b200; "
                                               "oracle/_ref/qsim_qtrajectory_refcuda = the reference's CUDA backend)")
ap.add_argument("--workers", type=int, default=1,
                help="worker threads per process (own state + CUDA per-thread stream each): one worker's host phases "
                     "overlap the other's kernels")
args = ap.parse_args()

with tempfile.NamedTemporaryFile("w", suffix=f"_rqc_q{args.n}", delete=False) as f:
    f.write(generate(args.n, args.depth, args.n))
    path = f.name
ppg = args.procs_per_gpu
total = args.total or args.num * args.gpus
extra = ("-b", str(args.batch), "-j", str(args.workers), "-x", str(args.prefix))
kw = {}
if args.binary:
    kw["binary"] = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), args.binary) \
        if not os.path.isabs(args.binary) else args.binary
    extra = ()   # the reference builds know neither batching nor prefix sharing
res = run_farm(path, 0, total, gpus=args.gpus * ppg, p=args.p, max_fused_size=args.fused,
               device_ids=[d for d in range(args.gpus) for _ in range(ppg)], extra_args=extra, **kw)
res["prefix_sharing"] = args.prefix if not args.binary else 0
res["binary"] = args.binary or "apps/_bin/qsim_qtrajectory_b200"
res["workers_per_gpu"] = args.workers
res["batch"] = args.batch
res["procs_per_gpu"] = ppg
os.unlink(path)
res.pop("sums")
res["mean"] = res["mean"][:8]
res.update({"config": f"rqc_q{args.n} depth {args.depth}, depolarize p={args.p}, f={args.fused}, "
                      f"{total} repetition ids over {args.gpus} GPU(s)", "total_repetitions_in_config": 8192,
            "est_full_8192_s": 8192 / res["trajectories_per_s"]})
print(json.dumps(res))
