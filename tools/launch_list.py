#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel -> small CSV for profiles/.
usage: python tools/launch_list.py raw.csv "command line that was profiled" > profiles/<tag>_launch_list.csv"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = OrderedDict()
for r in rows[1:]:
    a = agg.setdefault(r[k], [0, 0.0])
    a[0] += 1
    a[1] += float(r[v].replace(",", ""))
tot = sum(a[1] for a in agg.values())
print("# ncu launch list (gpu__time_duration.sum, --clock-control none) of:")
print("#   " + (sys.argv[2] if len(sys.argv) > 2 else "?"))
print(f"# {len(rows) - 1} launches, aggregated by kernel; per-launch times are cold-cache and serialised: compare SHARES")
print("kernel,launches,total_ns,share")
for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"\"{name[:90]}\",{n},{ns:.0f},{ns / tot:.4f}")
