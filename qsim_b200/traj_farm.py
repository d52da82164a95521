"""Trajectory farm: Monte-Carlo quantum trajectories split over the GPUs of one node.

Trajectories are independent units (lib/qtrajectory.h:268: the seed of a
trajectory is its repetition id), so the repetition ids [traj0, traj0 + num) are
block-partitioned over the GPUs, every GPU runs its slice in its own process
(apps/_bin/qsim_qtrajectory_b200 pinned with CUDA_VISIBLE_DEVICES) and the per-process
observable sums are added at the end: no collective on the data path, "replicas".
"""
import json
import os
import subprocess
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BINARY = os.path.join(ROOT, "apps", "_bin", "qsim_qtrajectory_b200")


def partition(traj0, num, parts):
    """Block partition of the repetition ids: [(first id, count)] per part, sizes differ by <= 1."""
    base, extra = divmod(num, parts)
    out, start = [], traj0
    for r in range(parts):
        cnt = base + (1 if r < extra else 0)
        out.append((start, cnt))
        start += cnt
    return out


def merge(results):
    """Adds the observable sums of the slices; rate = trajectories / slowest slice."""
    sums = None
    for r in results:
        s = r["sums"]
        sums = list(s) if sums is None else [a + b for a, b in zip(sums, s)]
    num = sum(r["num"] for r in results)
    slowest = max(r["seconds"] for r in results)
    n = results[0]["n"]
    gate_passes = sum(r.get("gate_passes", 0) for r in results)
    expect_passes = sum(r.get("expect_passes", 0) for r in results)
    # algorithmic bytes (fp32): gate pass 16*2^n, expectation pass 8*2^n, SetStateZero 8*2^n per trajectory
    # (the 3-4 read passes of a OneQubitMoments call are not counted: the figure drops when they replace
    # one pass per operator -- compare trajectories/s)
    abytes = (16.0 * gate_passes + 8.0 * expect_passes + 8.0 * num) * (1 << n)
    return {
        "n": n, "num": num, "slices": [(r["traj0"], r["num"], r["seconds"]) for r in results],
        "seconds": slowest, "trajectories_per_s": num / slowest if slowest > 0 else float("nan"),
        "gate_passes": gate_passes, "expect_passes": expect_passes,
        "moment_calls": sum(r.get("moment_calls", 0) for r in results),
        "prefix_gates_skipped": sum(r.get("prefix_gates_skipped", 0) for r in results),
        "noiseless_trajectories": sum(r.get("noiseless_trajectories", 0) for r in results),
        "kraus_group_hits": sum(r.get("kraus_group_hits", 0) for r in results),
        "algorithmic_GBps": abytes / slowest / 1e9 if slowest > 0 else float("nan"),
        "mean": [s / num for s in sums], "sums": sums,
    }


def run_farm(circuit_file, traj0, num, gpus=1, p=0.001, max_fused_size=4, maxtime=None,
             binary=BINARY, device_ids=None, extra_args=()):
    """Runs `num` trajectories of the depolarizing-noise version of `circuit_file` on `gpus` GPUs."""
    if not os.path.exists(binary):
        raise RuntimeError(f"{binary} is missing: run `python __graft_entry__.py` (build) first")
    device_ids = list(range(gpus)) if device_ids is None else list(device_ids)
    procs = []
    t0 = time.perf_counter()
    for dev, (start, cnt) in zip(device_ids, partition(traj0, num, gpus)):
        if cnt == 0:
            continue
        cmd = [binary, "-c", circuit_file, "-p", repr(p), "-0", str(start), "-n", str(cnt),
               "-f", str(max_fused_size)] + list(extra_args)
        if maxtime is not None:
            cmd += ["-d", str(maxtime)]
        env = dict(os.environ)
        if dev is not None:
            env["CUDA_VISIBLE_DEVICES"] = str(dev)
        procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env))
    results = []
    for pr in procs:
        out, err = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError(f"trajectory worker failed ({pr.returncode}): {err[-400:]}")
        results.append(json.loads(out.strip().splitlines()[-1]))
    merged = merge(results)
    merged["wall_s"] = time.perf_counter() - t0
    merged["gpus"] = len(procs)
    return merged
