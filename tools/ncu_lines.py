#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line.
usage: ncu_lines.py src.csv [file-substring] [topN]"""
import csv, sys
path = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cur_file = None; hdr = None; out = []
for r in csv.reader(open(path)):
    if not r: continue
    if r[0] in ("File Name", "File Path"): cur_file = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit(): continue
    if r[2] != "-": continue            # sass rows carry an address
    d = dict(zip(hdr[4:], r[4:]))
    try:
        inst = int(d["Instructions Executed"]); samp = int(d["# Samples"])
    except ValueError:
        continue
    if inst == 0 and samp == 0: continue
    stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "(Not" not in k and v.isdigit() and int(v) > 0}
    out.append((cur_file, int(r[0]), r[1].strip()[:70], inst, samp, stalls))
out = [o for o in out if want in (o[0] or "")]
ti = sum(o[3] for o in out); ts = sum(o[4] for o in out)
print(f"total inst {ti:,}  samples {ts:,}")
for key, name in ((3, "instructions"), (4, "samples")):
    print(f"--- top by {name}")
    for o in sorted(out, key=lambda o: -o[key])[:top]:
        st = ",".join(f"{k}:{v}" for k, v in sorted(o[5].items(), key=lambda kv: -kv[1])[:3])
        print(f"{o[1]:5d} inst {100*o[3]/max(ti,1):5.1f}% samp {100*o[4]/max(ts,1):5.1f}%  {o[2]:70s} {st}")
