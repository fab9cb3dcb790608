#!/usr/bin/env python
"""Summarise an `ncu --csv --metrics ...` launch list: mean of every metric per kernel name.
usage: ncu_launches.py launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[start]
ki, mi, vi, ui = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.defaultdict(lambda: collections.defaultdict(list))
units = {}
for r in rows[start + 1:]:
    if len(r) <= vi:
        continue
    try:
        agg[r[ki]][r[mi]].append(float(r[vi].replace(",", "")))
        units[r[mi]] = r[ui]
    except ValueError:
        pass
for k, ms in agg.items():
    name = k.split("(")[0].split("::")[-1]
    n = max(len(v) for v in ms.values())
    print(f"{name}  x{n}")
    for m, v in ms.items():
        if v:
            print(f"    {m:60s} {sum(v) / len(v):16.3f} {units[m]}")
