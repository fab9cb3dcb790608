#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by enclosing function of a source file.
usage: ncu_funcs.py src.csv path/to/source.cuh [file-substring]"""
import csv, re, sys
path, srcpath = sys.argv[1], sys.argv[2]
want = sys.argv[3] if len(sys.argv) > 3 else srcpath.split('/')[-1]
rows=[]; cur=None; hdr=None
for r in csv.reader(open(path)):
    if not r: continue
    if r[0] in ("File Name","File Path"): cur=r[1]; continue
    if r[0]=="Line No": hdr=r; continue
    if hdr is None or len(r)<len(hdr) or not r[0].isdigit(): continue
    if r[2]!="-": continue
    d=dict(zip(hdr[4:],r[4:]))
    try: inst=int(d["Instructions Executed"]); samp=int(d["# Samples"])
    except ValueError: continue
    if want in (cur or ''): rows.append((int(r[0]),inst,samp))
src=open(srcpath).read().split('\n')
marks=[]
for i,l in enumerate(src,1):
    m=re.search(r'__device__ (?:__forceinline__|__noinline__) \S+ (\w+)\(|__global__ void .* (\w+)\(|^struct (\w+)',l)
    if m: marks.append((i,[g for g in m.groups() if g][0]))
    elif re.search(r'// =+ (chain warp of|row warp of|producer warp$)', l): marks.append((i,'kernel:'+l.strip().strip('/= ')))
ti=sum(r[1] for r in rows); ts=sum(r[2] for r in rows)
agg={}
for ln,inst,samp in rows:
    name='?'
    for i,n in marks:
        if i<=ln: name=n
    a=agg.setdefault(name,[0,0]); a[0]+=inst; a[1]+=samp
print(f"total inst {ti:,} samples {ts:,}")
for n,(i,s) in sorted(agg.items(), key=lambda kv:-kv[1][0]):
    print(f"{n:40s} inst {100*i/ti:5.1f}% ({i/1e6:7.1f}M)  samp {100*s/ts:5.1f}%")
