#!/usr/bin/env python
"""apps/qsim_base_b200 -g N on the weak-scaling circuit rqc_q<n> (ONE process drives N GPUs through
StateSpaceB200Sharded / SimulatorB200Sharded / B200Runner): the printed amplitudes must equal the single-GPU
goldens (tests/golden/rqc_amplitudes.json).  usage: python tools/sp_check.py N [n]   (default n = 30 + log2 N)"""
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.gen_rqc import generate  # noqa: E402

shards = int(sys.argv[1])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30 + shards.bit_length() - 1
gold = json.load(open(os.path.join(ROOT, "tests", "golden", "rqc_amplitudes.json")))[f"rqc_q{n}_d20_f4"]
with tempfile.NamedTemporaryFile("w", suffix=f"_rqc_q{n}", delete=False) as f:
    f.write(generate(n, 20, n))
    path = f.name
t0 = time.time()
p = subprocess.run([os.path.join(ROOT, "apps", "_bin", "qsim_base_b200"), "-c", path, "-d", "20", "-f", "4", "-g", str(shards), "-v", "1"],
                   capture_output=True, text=True, timeout=900)
wall = time.time() - t0
os.unlink(path)
amps = {}
for line in p.stdout.splitlines():
    m = re.match(r"([01]{3}):\s+(\S+)\s+(\S+)\s+(\S+)", line)
    if m:
        amps[int(m.group(1), 2)] = complex(float(m.group(2)), float(m.group(3)))
known = {i: complex(*a) for i, a in zip(gold["indices"], gold["amplitudes"]) if i < 8}
err = max(abs(amps[i] - v) for i, v in known.items()) if amps else float("nan")
simu = re.search(r"simu time is ([0-9.eE+-]+) seconds", p.stdout + p.stderr)
ok = p.returncode == 0 and len(amps) == 8 and err < 5e-8
print(json.dumps({"shards": shards, "n": n, "mode": "single process, multi device (qsim_base_b200 -g)", "ok": bool(ok),
                  "max_abs_err_vs_single_gpu_golden": err, "amplitudes_compared": len(known),
                  "simu_time_s": float(simu.group(1)) if simu else None, "wall_s": wall, "stderr": p.stderr[-300:]}))
sys.exit(0 if ok else 1)
