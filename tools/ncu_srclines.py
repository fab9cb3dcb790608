#!/usr/bin/env python
"""Top source lines by executed warp instructions of one kernel in an ncu report.
usage: ncu_srclines.py report.ncu-rep kernel-regex [source-file-substring] [N]"""
import collections, csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
want = sys.argv[3] if len(sys.argv) > 3 else ""
N = int(sys.argv[4]) if len(sys.argv) > 4 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kre],
                     capture_output=True, text=True).stdout
hdr = None; cur = None
agg = collections.Counter()
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] in ("File Name", "File Path"): cur = r[1] if len(r) > 1 else None; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or not r[0].isdigit(): continue
    d = dict(zip(hdr, r))
    try: inst = int(d.get("Instructions Executed", "0") or 0)
    except ValueError: continue
    agg[(cur or "", int(r[0]))] += inst
tot = sum(agg.values())
print(f"total {tot / 1e6:.1f}M warp instructions")
cache = {}
for (f, ln), v in agg.most_common(N):
    if want and want not in f: continue
    try:
        if f not in cache: cache[f] = open(f.replace("/tmp/code/gotaku6629__CTC/repo", "/root/repo")).read().split("\n")
        text = cache[f][ln - 1].strip()[:110]
    except Exception: text = ""
    print(f"{v / 1e6:8.1f}M {100 * v / tot:5.1f}%  {f.split('/')[-1]}:{ln}  {text}")
