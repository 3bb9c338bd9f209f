#!/usr/bin/env python
"""Per-kernel summary of an ncu launch list (--metrics gpu__time_duration.sum --csv): launches, mean / max duration, share
of the serialised kernel time.   python profiles/summarise_launches.py gpurun_out/launches.csv [skip_first_n]"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
    rows.append((re.sub(r"\(.*", "", r["Kernel Name"]).replace("plviwo::", ""), v))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
agg = defaultdict(list)
for k, v in rows:
    agg[k].append(v)
tot = sum(v for _, v in rows)
print("%-44s %7s %10s %10s %10s %7s" % ("kernel", "n", "mean us", "max us", "sum us", "share"))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%-44s %7d %10.1f %10.1f %10.1f %6.1f%%" % (k[:44], len(v), sum(v) / len(v), max(v), sum(v), 100 * sum(v) / tot))
print("%-44s %7d %10s %10s %10.1f" % ("total", len(rows), "", "", tot))
