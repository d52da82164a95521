#!/usr/bin/env python
"""bench.py -- the reference's headline benchmark on B200.

Workload (BASELINE.json configs[1]): circuits/circuit_q30, depth 20, fused by the
reference's own MultiQubitGateFuser with max_fused_size=4 -> 41 fused-gate passes
over a 2^30-amplitude fp32 state (8 GiB).  The fused-gate trace is the golden
fixture tests/golden/q30_d20_f4.trace written by oracle/ref_fuse.cc.

A "step" = the 41 passes on a resident state (|0...0> re-initialised outside the
timed region).  metric = algorithmic fused-gate HBM GB/s = 41 * 16 * 2^30 B / time.
e2e = the same through the public API from host buffers: SetStateZero + 41
ApplyGate calls with host matrices + 8 GetAmpl + Norm, wall clock.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TRACE = os.path.join(ROOT, "tests", "golden", "q30_d20_f4.trace")
METRIC = "rqc_q30_d20_f4_fused_gate_hbm_gbs"
WORKLOAD = ("circuits/circuit_q30 depth 20, max_fused_size 4: 41 fused-gate passes (reference parser + fuser output, "
            "tests/golden/q30_d20_f4.trace) on a 2^30-amplitude fp32 state (8 GiB)")


# stdout carries exactly ONE JSON line: everything else a library prints there (NCCL's version banner,
# torchrun notices) is sent to stderr by pointing fd 1 at fd 2 for the whole run.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def sustained_copy_gbs(seconds=1.5):
    """Copy bandwidth of THIS box in steady state, for context beside MEASURED_PEAKS.json (a best-of-10 burst): a
    2 GiB device-to-device torch copy repeated back to back for ~1.5 s (read + write bytes, CUDA events) -- what a
    plain copy sustains once the board sits at its power cap, like the 41 passes of a step do."""
    try:
        import torch
        a = torch.empty(1 << 29, dtype=torch.float32, device="cuda")
        b = torch.empty_like(a)
        for _ in range(3):
            b.copy_(a)
        torch.cuda.synchronize()
        reps = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        while True:
            for _ in range(20):
                b.copy_(a)
            reps += 20
            torch.cuda.synchronize()
            if time.perf_counter() - t0 > seconds:
                break
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        del a, b
        return 2.0 * 4 * (1 << 29) * reps / (ms * 1e-3) / 1e9
    except Exception:
        return None


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(ops, n, min_seconds=10.0, max_gates=None, threads=None):
    """Times the reference's own CPU path (oracle/_ref: SimulatorAVX512/AVX via
    lib/simmux.h, OpenMP) on a bounded sample: the first gates of the same trace on
    the same 2^n state, until `min_seconds` of gate time has elapsed."""
    from oracle.oracle import SIMD_F32, RefEngine, ref_library_path
    threads = threads or os.cpu_count() or 1
    kind = "reference" if ref_library_path() else "port"
    algo_bytes = 16.0 * (1 << n)
    if kind == "reference":
        eng = RefEngine(SIMD_F32, n, threads)
        eng.set_zero()
        name = eng.simd_name()
        apply = lambda op: eng.apply_gate(op.qubits, op.matrix)
    else:
        from oracle.oracle import Oracle
        orc = Oracle()
        st = np.zeros(1 << n, np.complex64); st[0] = 1
        name = "oracle port (plain C + OpenMP)"
        apply = lambda op: orc.apply_gate(st, op.qubits, op.matrix)
    t_total, done = 0.0, 0
    for op in ops[: max_gates or len(ops)]:
        t0 = time.perf_counter()
        apply(op)
        t_total += time.perf_counter() - t0
        done += 1
        if t_total >= min_seconds and done >= 4:
            break
    gbs = done * algo_bytes / t_total / 1e9
    return {"value": gbs, "unit": "GB/s", "cores": threads, "kind": kind,
            "sample": f"first {done} of {len(ops)} fused gates of circuit_q{n} d20 f4 on the full 2^{n} fp32 state, "
                      f"{name}, {threads} OpenMP threads, {t_total:.1f} s",
            "seconds": t_total, "gates": done,
            "est_full_circuit_s": t_total / done * len(ops)}


def gpu_reference_baseline(fused=4, runs=2):
    """The reference's OWN CUDA backend (apps/qsim_base_cuda.cu recompiled for sm_100a by oracle/Makefile,
    oracle/_ref/qsim_base_cuda_ref) on the same GPU and circuit: the GPU kernel-to-beat of SURVEY 2.3.  Its
    own "-v 2" clock ("simu time": the fused-gate loop bracketed by device synchronisation)."""
    import re
    exe = os.path.join(ROOT, "oracle", "_ref", "qsim_base_cuda_ref")
    circ = os.path.join(ROOT, "oracle", "_ref", "circuits", "circuit_q30")
    if not (os.path.exists(exe) and os.path.exists(circ)):
        return {"kind": "unavailable", "why": "oracle/_ref/qsim_base_cuda_ref not built (reference tree absent at build time)"}
    times, amp0 = [], None
    try:
        for _ in range(runs + 1):
            p = subprocess.run([exe, "-c", circ, "-d", "20", "-f", str(fused), "-v", "2"], capture_output=True, text=True, timeout=300)
            m = re.search(r"simu time is ([0-9.eE+-]+) seconds", p.stdout + p.stderr)
            a = re.search(r"^000:\s+(\S+)\s+(\S+)", p.stdout, flags=re.M)
            if p.returncode != 0 or not m:
                return {"kind": "unavailable", "why": (p.stderr or p.stdout)[-300:]}
            times.append(float(m.group(1)))
            amp0 = [float(a.group(1)), float(a.group(2))] if a else None
    except Exception as e:
        return {"kind": "unavailable", "why": str(e)}
    best = min(times[1:])   # first run pays context creation inside the process but outside "simu time"; keep it out anyway
    return {"kind": "reference CUDA backend (lib/simulator_cuda.h) recompiled for sm_100a, same GPU, same circuit, -f %d" % fused,
            "ms_per_circuit": best * 1e3, "runs_ms": [t * 1e3 for t in times], "amp0": amp0,
            "value": 41 * 16.0 * (1 << 30) / best / 1e9 if fused == 4 else None, "unit": "GB/s"}


Q30_KNOWN = {0: (1.4871957e-5, 2.8161678e-5), 1: (1.8767701e-5, 7.3190154e-6),
             2: (-1.1130518e-5, 1.6207156e-5), 7: (-1.622395e-5, 3.5199686e-5)}   # reference qsim_base, BASELINE.md 4


def parity_block(get_ampl, norm, known, tol, what):
    errs = [abs(get_ampl(i) - complex(*v)) for i, v in known.items()]
    ok = bool(max(errs) <= tol and abs(norm - 1.0) < 1e-4)
    return {"ok": ok, "amplitudes_checked": len(errs), "max_abs_err": float(max(errs)), "tolerance": tol,
            "norm": norm, "against": what}


def run_reference_arm(args, rank):
    import qsim_b200
    if rank != 0:
        return
    n, ops = qsim_b200.read_trace(TRACE)
    per_step = []
    res = None
    # a step = ALL fused gates of the circuit (same config as the GPU arm); ~5 s per step on 16-32 host cores
    for it in range(args.warmup + args.steps):
        res = cpu_reference_run(ops, n, min_seconds=1e9)
        if it >= args.warmup:
            per_step.append(res)
    val = float(np.mean([r["value"] for r in per_step]))
    ms = float(np.mean([r["seconds"] for r in per_step])) * 1e3
    res = dict(per_step[-1]); res["value"] = val
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "GB/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "l2": "state (8 GiB) far larger than any cache",
                       "arm": "the reference's own CPU path (SimulatorAVX512/AVX via oracle/_ref, all host cores), every one of "
                              "the 41 fused gates per step"},
            "cpu_baseline": res,
            "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def run_sharded(args, rank, world, local_rank, dist):
    """N > 1: weak scaling.  One 2^30-amplitude shard (8 GiB) per GPU, n = 30 + log2(N) qubits, circuit built like
    circuit_q30 by tools/gen_rqc.py and fused by the reference fuser (tests/golden/rqc_q<n>_d20_f4.trace).  The
    state is a multi-process qb200_sv (csrc/sharded.cu through qsim_b200/sv.py): the library plans the
    local<->global exchanges over the fused-gate list and runs them as one push kernel per GPU over NVLink peer
    memory; torch.distributed only carries the three host-side collectives of qb200_comm."""
    import torch
    import qsim_b200
    from qsim_b200 import _lib
    from qsim_b200.sv import ShardedStateB200, pack_gates

    g = world.bit_length() - 1
    n = args.shard_qubits + g
    trace = os.path.join(ROOT, "tests", "golden", f"rqc_q{n}_d20_f4.trace")
    nq, ops = qsim_b200.read_trace(trace)
    assert nq == n
    host_group = dist.new_group(backend="gloo")   # host-side collectives of qb200_comm: a few doubles, CPU tensors
    sv = ShardedStateB200.multi_process(dist, n, local_rank, host_group=host_group)
    for kv in args.tune:
        key, val = kv.split("=")
        sv.set_option(key, int(val))
    packed = pack_gates(ops)
    lib = _lib.load()
    ctx0 = sv.shards()[0][3]

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    def one_run():
        sv.SetStateZero(reset_map=True)
        sv.Run(packed=packed)

    for _ in range(max(args.warmup, 3)):
        one_run()
    barrier()
    sv.reset_stats()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = sv.launch_count()
    dev_ms = 0.0
    ms = C_float()
    for _ in range(args.steps):
        sv.SetStateZero(reset_map=True)
        barrier()
        lib.qb200_timer_start(ctx0)
        sv.Run(packed=packed)
        lib.qb200_timer_stop_ms(ctx0, ms)
        dev_ms += float(ms.value)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = sv.launch_count() - l0
    stats = sv.stats()
    t = torch.tensor([dev_ms, stats["exchange_ms"], stats["overlap_ms"]], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, exch_ms, ovl_ms = float(t[0].item()), float(t[1].item()), float(t[2].item())
    ms_per_step = dev_ms / args.steps
    total_bytes = len(ops) * 16.0 * (1 << n)
    value = total_bytes / (ms_per_step * 1e-3) / 1e9

    # ---- end to end from host buffers (wall clock): plan (cached after the first call) + gates + 8 amplitudes + norm
    e2e_ms = []
    amps, nrm = None, None
    phases = {"issue_ms": 0.0, "amplitudes_ms": 0.0, "norm_ms": 0.0}
    for it in range(1 + args.steps):
        barrier()
        t0 = time.perf_counter()
        one_run()
        t1 = time.perf_counter()
        amps = sv.GetAmpls(range(8))   # waits for the circuit; one host-side collective
        t2 = time.perf_counter()
        nrm = sv.Norm()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        if it >= 1:
            e2e_ms.append((t3 - t0) * 1e3)
            phases["issue_ms"] += (t1 - t0) * 1e3 / args.steps
            phases["amplitudes_ms"] += (t2 - t1) * 1e3 / args.steps
            phases["norm_ms"] += (t3 - t2) * 1e3 / args.steps
    t = torch.tensor([float(np.mean(e2e_ms))], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_step = float(t.item())

    # ---- parity: 64 amplitudes + norm against the single-GPU goldens (tools/make_rqc_goldens.py) ----
    # (config 4, 128 GiB shards: no single GPU holds the state; norm == 1 is what can be checked at that size,
    #  the same code path is compared with goldens at 8 GiB shards)
    parity = {"ok": bool(abs(nrm - 1.0) < 1e-4), "amplitudes_checked": 0, "norm": nrm,
              "why": "no golden for this circuit size: norm only"}
    try:
        with open(os.path.join(ROOT, "tests", "golden", "rqc_amplitudes.json")) as f:
            gold = json.load(f).get(f"rqc_q{n}_d20_f4")
        if gold and args.shard_qubits == 30:
            known = {int(i): tuple(a) for i, a in zip(gold["indices"], gold["amplitudes"])}
            vals = dict(zip(known, sv.GetAmpls(list(known))))
            parity = parity_block(vals.get, nrm, known, 5e-8,
                                  "tests/golden/rqc_amplitudes.json: the same circuit on ONE GPU through the single-GPU path "
                                  "(q31 also equal to the reference AVX-512 simulator to 1e-10)")
            parity["golden_norm"] = gold["norm"]
    except Exception as e:
        parity = {"ok": False, "why": str(e)}
    if rank == 0:
        peaks, peak_kind = measured_peaks()
        swaps = int(stats["swaps"]) // args.steps
        sent = stats["bytes_sent_per_shard"] / args.steps
        line = {"metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"rqc_q{n} depth 20 (tools/gen_rqc.py, circuit_q30 rules), max_fused_size 4: {len(ops)} fused-gate "
                                       f"passes on a 2^{n}-amplitude fp32 state sharded over {world} GPUs ({8 << (n - g - 30)} GiB shard each)",
                           "l2": f"shard ({8 << (n - g - 30)} GiB) is far larger than L2",
                           "multi_gpu": f"global-qubit sharding behind the C ABI (qb200_sv_*): {swaps} local<->global exchanges per circuit "
                                        f"planned over the fused-gate list with commuting gates reordered (csrc/sv_plan.h), each ONE push "
                                        f"kernel per GPU over NVLink peer memory; {int(stats['local_swap_passes']) // args.steps} local SWAP passes",
                           "wall_time_s_per_circuit": ms_per_step * 1e-3},
                "roofline": {"bound": "hbm", "achieved": value / world, "peak": float(peaks["hbm_gbs"]), "unit": "GB/s",
                             "frac": value / world / float(peaks["hbm_gbs"]), "traffic": None,
                             "note": "per-GPU algorithmic gate bytes over the whole step (exchange time included)",
                             "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind})"},
                "swap": {"swaps_per_circuit": swaps, "bytes_sent_per_rank_per_circuit": sent,
                         "exchange_ms_per_circuit": exch_ms / args.steps,
                         "barrier_wait_ms_per_circuit_rank0": stats["barrier_wait_ms"] / args.steps,
                         "overlapped_exchanges_per_circuit": int(stats["overlapped_swaps"]) // args.steps,
                         "gate_passes_pipelined_against_them": int(stats["overlapped_gate_passes"]) // args.steps,
                         "overlapped_pipeline_ms_per_circuit": ovl_ms / args.steps,
                         "note": "an overlapped exchange runs chunk by chunk on a second stream beside the last gate passes of its "
                                 "epoch (csrc/sharded.cu run_overlapped); its time is inside overlapped_pipeline_ms (gates included), "
                                 "exchange_ms covers the exchanges that ran alone",
                         "nvlink_GBps_per_direction": (sent / (exch_ms / args.steps * 1e-3) / 1e9
                                                       if exch_ms > 0 and not stats["overlapped_swaps"] else None),
                         "nvlink_peak_GBps_per_direction": 900.0},
                "parity": parity,
                "cpu_baseline": None,
                "e2e": {"value": total_bytes / (e2e_step * 1e-3) / 1e9, "unit": "GB/s",
                        "h2d_bytes_per_step": sum(op.matrix.nbytes for op in ops), "d2h_bytes_per_step": 72,
                        "ms_per_step": e2e_step, "amp0": [amps[0].real, amps[0].imag], "norm": nrm,
                        "host_phases_rank0": phases},
                "gpu_launches": int(launches), "clocks": clocks}
        emit(line)
    sv.close()
    return parity.get("ok", False)


def C_float():
    import ctypes
    return ctypes.c_float()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--trace", default=TRACE)
    ap.add_argument("--tune", action="append", default=[], help="key=value tuning override (N = 1), e.g. tc=0")
    ap.add_argument("--shard-qubits", type=int, default=30,
                    help="N > 1 only: qubits per shard (30 = 8 GiB, the driver's weak-scaling point; "
                         "34 = 128 GiB: BASELINE config 4, 36 qubits on 4 GPUs / 37 on 8)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import qsim_b200

    torch.cuda.set_device(local_rank)
    # NCCL_DEBUG is left as the caller set it: fd 1 already points at stderr (top of this file), so INFO lines
    # cannot reach the JSON line
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if world > 1:
        ok = run_sharded(args, rank, world, local_rank, dist)
        dist.destroy_process_group()
        if not ok:
            raise SystemExit("parity check of the sharded state FAILED (see the parity block of the JSON line)")
        return

    n, ops = qsim_b200.read_trace(args.trace)
    ss = qsim_b200.StateSpaceB200(np.float32, device=local_rank)
    sim = qsim_b200.SimulatorB200(np.float32, device=local_rank)
    for kv in args.tune:
        key, val = kv.split("=")
        sim.set_tuning(key, int(val))
    st = ss.Create(n)
    if ss.IsNull(st):
        raise SystemExit("not enough device memory for the state")
    amps = 1 << n
    pass_bytes = 16.0 * amps

    def apply_all():
        for op in ops:
            if op.controls:
                sim.ApplyControlledGate(op.qubits, op.controls, op.cvals, op.matrix, st)
            else:
                sim.ApplyGate(op.qubits, op.matrix, st)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-timed steps (state resident, CUDA events on the launch stream) ----
    for _ in range(max(args.warmup, 3)):
        ss.SetStateZero(st)
        apply_all()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = sim.launch_count()
    step_ms = []
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        ss.SetStateZero(st)
        ss.DeviceSync()
        sim.timer_start()
        apply_all()
        step_ms.append(sim.timer_stop_ms())
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = sim.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = float(np.sum(step_ms))
    if dist is not None:
        t = torch.tensor([dev_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps
    total_bytes = world * len(ops) * pass_bytes
    value = total_bytes / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the public API from host buffers (wall clock) ----
    e2e_ms = []
    h2d = sum(op.matrix.nbytes + 4 * (len(op.qubits) + len(op.controls)) for op in ops)
    d2h = 8 * 8 + 8
    amp_out = None
    for it in range(2 + args.steps):
        barrier()
        t0 = time.perf_counter()
        ss.SetStateZero(st)
        apply_all()
        amp_out = [ss.GetAmpl(st, i) for i in range(8)]
        nrm = ss.Norm(st)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if it >= 2:
            e2e_ms.append(dt * 1e3)
    e2e_step = float(np.mean(e2e_ms))
    if dist is not None:
        t = torch.tensor([e2e_step], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_step = float(t.item())
    e2e_value = total_bytes / (e2e_step * 1e-3) / 1e9

    # ---- per-launch durations, grouped by the kernel the dispatcher picked (named by the library itself:
    # qb200_last_kernel_name, set in gate_launch.cuh where the choice is made) ----
    ss.SetStateZero(st)
    per_gate = []
    for op in ops:
        sim.timer_start()
        sim.ApplyGate(op.qubits, op.matrix, st)
        per_gate.append((len(op.qubits), list(op.qubits), sim.timer_stop_ms(), sim.last_kernel_name()))
    kind = {"k_gate_tca": "tcgen05 3xTF32, A in TMEM", "k_gate_tc": "tcgen05 3xTF32, operands in smem",
            "k_gate_tcx": "tcgen05 3xTF32, A in TMEM", "k_gate_tile": "FFMA2, warp tile",
            "k_gate_pipe": "FFMA2, cp.async ring", "k_gate_reg": "FFMA2, registers", "k_gate_big": "FFMA2, row blocks",
            "k_gate_tcl": "tcgen05 3xTF32, low targets staged through shared memory"}

    def kernel_class(g, qs, name):
        return f"{name} ({kind.get(name.split('<')[0], 'generic')})"

    classes = {}
    for g, qs, ms, name in per_gate:
        classes.setdefault(kernel_class(g, qs, name), []).append(ms)
    total_ms = float(np.sum([p[2] for p in per_gate]))
    kernels = {k: {"launches": len(v), "avg_ms": float(np.mean(v)), "GBps": pass_bytes / (float(np.mean(v)) * 1e-3) / 1e9,
                   "share_of_step": float(np.sum(v)) / total_ms} for k, v in classes.items()}
    dom = max(kernels, key=lambda k: kernels[k]["share_of_step"])
    peaks, peak_kind = measured_peaks()
    peak = float(peaks["hbm_gbs"])
    dom_ms = kernels[dom]["avg_ms"]
    achieved = pass_bytes / (dom_ms * 1e-3) / 1e9
    # DRAM bytes per launch of the same kernel from the committed `ncu --set full` capture
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")) as f:
            tj = json.load(f)
        key = next(k for k in tj["kernels"] if dom.startswith(k.split("<")[0] + "<") and k.split("<")[1][0] == dom.split("<")[1][0])
        k = tj["kernels"][key]
        traffic = (k["dram_bytes_read"] + k["dram_bytes_write"]) * (amps / float(1 << 30))
        traffic_src = f"profiles/r02_ncu_traffic.json [{key}] (dram__bytes_read.sum + dram__bytes_write.sum per launch)"
    except Exception:
        pass
    sustained = sustained_copy_gbs()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "sustained_copy_GBps_this_box": sustained,
                "frac_of_sustained_copy": achieved / sustained if sustained else None,
                "whole_step_frac": value / peak,
                "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": pass_bytes,
                "kernel": dom, "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind})",
                "avg_launch_ms": dom_ms, "launches_timed": kernels[dom]["launches"],
                "share_of_step": kernels[dom]["share_of_step"], "kernels": kernels,
                "per_gate_ms": [round(p[2], 4) for p in per_gate],
                "per_gate_lowest_target": [p[1][0] if p[1] else None for p in per_gate]}

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            try:
                cpu = cpu_reference_run(ops, n, min_seconds=1e9)   # the whole circuit, like the GPU arm
            except Exception as e:  # keep the bench line even if the host lacks RAM for the sample
                cpu = {"value": None, "unit": "GB/s", "cores": os.cpu_count(), "kind": "unavailable", "sample": str(e)}
        parity = parity_block(lambda i: amp_out[i] if i < 8 else ss.GetAmpl(st, i), nrm, Q30_KNOWN, 2e-9,
                              "amplitudes printed by the reference's apps/qsim_base.cc (AVX-512) for circuit_q30 -d 20 "
                              "(BASELINE.md section 4); full-state comparison: tests/test_circuit_gpu.py") \
            if os.path.basename(args.trace).startswith("q30_d20") else None
        # second configuration BASELINE configs[1] names: the same circuit fused to 5 qubits (44 passes)
        f5 = None
        f5_trace = os.path.join(ROOT, "tests", "golden", "q30_d20_f5.trace")
        if args.trace == TRACE and os.path.exists(f5_trace):
            n5, ops5 = qsim_b200.read_trace(f5_trace)
            t5 = []
            for it in range(2 + args.steps):
                ss.SetStateZero(st)
                ss.DeviceSync()
                sim.timer_start()
                for op in ops5:
                    sim.ApplyGate(op.qubits, op.matrix, st)
                if it >= 2:
                    t5.append(sim.timer_stop_ms())
                else:
                    sim.timer_stop_ms()
            a5 = [ss.GetAmpl(st, i) for i in Q30_KNOWN]
            f5 = {"workload": "circuit_q30 depth 20, max_fused_size 5: %d passes" % len(ops5), "ms_per_step": float(np.mean(t5)),
                  "value": len(ops5) * pass_bytes / (float(np.mean(t5)) * 1e-3) / 1e9, "unit": "GB/s",
                  "max_abs_err_vs_reference_amplitudes": float(max(abs(a - complex(*v)) for a, v in zip(a5, Q30_KNOWN.values())))}
        del st
        gpu_base = None
        if not args.no_cpu_baseline and world == 1 and args.trace == TRACE:
            gpu_base = gpu_reference_baseline(4)
            if gpu_base.get("ms_per_circuit"):
                gpu_base["ours_over_reference_cuda"] = gpu_base["ms_per_circuit"] / ms_per_step
        line = {"metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD if args.trace == TRACE else
                           f"{os.path.basename(args.trace)}: {len(ops)} fused-gate passes on a 2^{n}-amplitude fp32 state",
                           "l2": "state (8 GiB) is 68x larger than L2: every pass streams from HBM",
                           "multi_gpu": "independent replicas" if world > 1 else "single GPU",
                           "wall_time_s_per_circuit": ms_per_step * 1e-3},
                "roofline": roofline, "cpu_baseline": cpu, "gpu_baseline": gpu_base, "fused5": f5, "parity": parity,
                "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_step, "amp0": [amp_out[0].real, amp_out[0].imag], "norm": nrm},
                "gpu_launches": int(launches), "clocks": clocks,
                "timed_region_wall_s": t_wall}
        emit(line)


if __name__ == "__main__":
    main()
