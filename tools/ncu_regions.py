#!/usr/bin/env python
"""Per-source-line and per-function-region instruction / sample shares of an `ncu --page source --csv` dump.
usage: ncu_regions.py src.csv source-file [topN]"""
import csv, re, sys
path, srcpath = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
src = open(srcpath).read().split("\n")
base = srcpath.split("/")[-1]
cur = None; hdr = None; out = {}
for r in csv.reader(open(path)):
    if not r: continue
    if r[0] in ("File Name", "File Path"): cur = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit(): continue
    if r[2] != "-" or base not in (cur or ""): continue
    d = dict(zip(hdr[4:], r[4:]))
    try: inst = int(d["Instructions Executed"]); samp = int(d["# Samples"])
    except ValueError: continue
    st = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "(Not" not in k and v.isdigit() and int(v) > 0}
    out[int(r[0])] = (inst, samp, st)
ti = sum(v[0] for v in out.values()); ts = sum(v[1] for v in out.values())
print(f"total inst {ti:,} samples {ts:,}")
# regions = enclosing function-like headers (lines starting a __device__/template function or a lambda 'auto x = [&]')
heads = [(i + 1, l.strip()[:70]) for i, l in enumerate(src) if re.match(r"\s*(__device__|__global__|static __device__|auto \w+ = \[&\])", l)]
def region(ln):
    name = "?"
    for h, n in heads:
        if h <= ln: name = f"{h}:{n}"
        else: break
    return name
agg = {}
for ln, (i, s_, st) in out.items():
    a = agg.setdefault(region(ln), [0, 0]); a[0] += i; a[1] += s_
print("--- regions")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"  inst {100*v[0]/ti:5.1f}% samp {100*v[1]/ts:5.1f}%  {k}")
print("--- lines by samples")
for ln, v in sorted(out.items(), key=lambda kv: -kv[1][1])[:top]:
    st = ",".join(f"{k}:{x}" for k, x in sorted(v[2].items(), key=lambda kv: -kv[1])[:3])
    print(f"  {ln:5d} inst {100*v[0]/ti:5.1f}% samp {100*v[1]/ts:5.1f}%  {src[ln-1].strip()[:80]:80s} {st}")
