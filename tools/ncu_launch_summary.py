#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: tools/ncu_launch_summary.py gpurun_out/launches.csv "command line that was profiled" > profiles/rNN_....txt"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[ik]).replace("qb200::", "")
    ns = float(r[iv].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6}.get(r[iu], 1)
    c = agg.setdefault(name, [0, 0.0])
    c[0] += 1
    c[1] += ns
total = sum(v[1] for v in agg.values())
print(f"ncu --metrics gpu__time_duration.sum --clock-control none: {sys.argv[2] if len(sys.argv) > 2 else ''}")
print("per-launch times under ncu are serialised/cold: compare SHARES with bench.py's share_of_step")
print(f"{'kernel':95s} {'count':>6s} {'total ms':>10s} {'avg ms':>8s} {'share':>7s}")
for name, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:95]:95s} {c:6d} {ns / 1e6:10.3f} {ns / 1e6 / c:8.3f} {100 * ns / total:6.1f}%")
