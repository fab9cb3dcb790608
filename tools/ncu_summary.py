#!/usr/bin/env python
"""Summary of one kernel of an ncu --set full report as JSON (for profiles/): duration, DRAM bytes, occupancy, issue
rate, stall reasons per issued instruction, cache hit rates and the SASS instructions with the most stall samples.
usage: ncu_summary.py report.ncu-rep [kernel-regex] > profiles/<name>.json"""
import csv
import io
import json
import re
import subprocess
import sys

rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else "."


def page(name, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


raw = page("raw")
hdr, units = raw[0], raw[1]
row = next(r for r in raw[2:] if re.search(kre, r[hdr.index("Kernel Name")]))
val = {k: (row[i], units[i]) for i, k in enumerate(hdr)}


def num(k):
    try:
        return float(val[k][0].replace(",", ""))
    except (KeyError, ValueError):
        return None


def scaled(k):  # bytes / time metrics come with a unit prefix
    v = num(k)
    if v is None:
        return None
    u = val[k][1]
    mul = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1}.get(u.split("/")[0], 1)
    return v * mul


out = {
    "report": rep.split("/")[-1],
    "kernel": val["Kernel Name"][0],
    "grid": val["Grid Size"][0], "block": val["Block Size"][0],
    "duration_us": (scaled("gpu__time_duration.sum") or 0) * 1e6,
    "dram_bytes_read": scaled("dram__bytes_read.sum"), "dram_bytes_write": scaled("dram__bytes_write.sum"),
    "registers_per_thread": num("launch__registers_per_thread"),
    "shared_mem_per_block_bytes": scaled("launch__shared_mem_per_block"),
    "occupancy_limit_blocks": {k.split("limit_")[1]: num(k) for k in hdr if k.startswith("launch__occupancy_limit_")},
    "sm__warps_active_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "smsp__issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "warp_instructions": num("smsp__inst_executed.sum"),
    "l1_hit_rate_pct": num("l1tex__t_sector_hit_rate.pct"), "l2_hit_rate_pct": num("lts__t_sector_hit_rate.pct"),
    "dram_cycles_active_pct": num("dram__cycles_active.avg.pct_of_peak_sustained_elapsed"),
    "stall_per_issue": {k.split("stalled_")[1].split("_per_")[0]: round(num(k), 3) for k in hdr
                        if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("per_issue_active.ratio") and (num(k) or 0) >= 0.05},
}
if out["dram_bytes_read"] is not None and out["duration_us"]:
    out["dram_GBps"] = (out["dram_bytes_read"] + out["dram_bytes_write"]) / out["duration_us"] / 1e3
src = page("source", ("--print-source", "sass", "-k", "regex:" + kre))
h = next((r for r in src if r and r[0] == "Address"), None)
if h:
    ix = {k: i for i, k in enumerate(h)}
    data = [r for r in src[src.index(h) + 1:] if len(r) == len(h)]
    tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
    top = sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:12]
    out["top_stall_instructions"] = [
        {"sass": r[ix["Source"]].strip(), "samples_pct": round(100 * int(r[ix["# Samples"]]) / tot, 1),
         "main_stall": max(((k, int(r[i])) for k, i in ix.items() if k.startswith("stall_") and "Not" not in k and r[i].isdigit()),
                           key=lambda kv: kv[1])[0]} for r in top]
print(json.dumps(out, indent=1))
