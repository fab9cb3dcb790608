#!/usr/bin/env python
"""SASS opcode histogram of the hot kernel instances in ctc_b200/libnbctc.so (cuobjdump -sass), for profiles/:
proves what the kernels are made of (TMA bulk copies UBLKCP / UBLKPF, mbarrier SYNCS, async copies LDGSTS, warp
reductions CREDUX, float64 DFMA/DMUL, no tensor-core UTC*MMA / HMMA on this path).
usage: sass_histogram.py [regex ...] > profiles/rNN_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ctc_b200", "libnbctc.so")
pats = sys.argv[1:] or [r"seqwarp_kernelILi1ELi5E", r"seqwarp_kernelILi2ELi5E", r"seqwide_kernelILi8ELi8ELb1E",
                        r"nbctc_stream_kernelILi2ELi8ELi5ELi1E", r"bin_emis_kernelILi5E", r"lattice_tile_kernelILi2E",
                        r"bin_grad_kernelILi5ELi32E", r"aux_ce_kernel", r"seqwarp_prep_kernel"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
cur, hist = None, collections.OrderedDict()
for line in out.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
    if m and cur:
        t = m.group(1).split()
        op = t[1] if t[0].startswith("@") else t[0]
        hist.setdefault(cur, collections.Counter())[op.split(".")[0]] += 1
for pat in pats:
    for fn, c in hist.items():
        if re.search(pat, fn):
            demangled = subprocess.run(["cu++filt", fn], capture_output=True, text=True).stdout.strip() or fn
            tot = sum(c.values())
            print(f"== {demangled}: {tot} SASS instructions")
            print("   " + ", ".join(f"{k} {v}" for k, v in c.most_common(28)))
            marks = {k: c.get(k, 0) for k in ("UBLKCP", "UBLKPF", "SYNCS", "LDGSTS", "CREDUX", "DFMA", "DMUL", "MUFU", "SHFL", "CCTL",
                                               "UTCHMMA", "UTCQMMA", "HMMA", "ATOM", "RED")}
            print("   markers: " + ", ".join(f"{k}={v}" for k, v in marks.items()))
