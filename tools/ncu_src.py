#!/usr/bin/env python
"""Summarise the ncu source page (SASS view) of a kernel by code REGION: instructions executed and stall samples
between marker instructions.  usage: ncu_src.py report.ncu-rep [n_top]"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# one block per kernel: a "Kernel Name" row, a header row, then one row per instruction
blocks, cur_rows = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur_rows = [r]
        blocks.append(cur_rows)
    elif cur_rows is not None:
        cur_rows.append(r)
which = sys.argv[3] if len(sys.argv) > 3 else "pileup_count"
blk = [b for b in blocks if which in b[0][1]][-1]
print(blk[0][1])
hdr = blk[1]
ia, isrc, isamp, iex, ithr = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
data = [r for r in blk[2:] if len(r) == len(hdr)]
tot_ex = sum(int(r[iex] or 0) for r in data)
tot_s = sum(int(r[isamp] or 0) for r in data)
print("instructions", len(data), "executed", tot_ex, "samples", tot_s)
# contiguous regions with the same execution count magnitude: split where the count changes by > 30 %
regions = []
cur = None
for k, r in enumerate(data):
    ex = int(r[iex] or 0)
    if cur is None or not (0.77 * cur["ex0"] <= ex <= 1.3 * cur["ex0"]):
        cur = dict(start=k, ex0=max(ex, 1), n=0, ex=0, samp=0, thr=0, atoms=0, first=r[isrc].strip())
        regions.append(cur)
    cur["n"] += 1
    cur["ex"] += ex
    cur["samp"] += int(r[isamp] or 0)
    cur["thr"] += int(r[ithr] or 0)
    if "ATOMS" in r[isrc]:
        cur["atoms"] += 1
regions.sort(key=lambda x: -x["ex"])
for g in regions[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print("  @%5d n=%4d exec/instr=%10d share=%5.1f%% stall-share=%5.1f%% lanes=%4.1f atoms=%3d  %s" % (
        g["start"], g["n"], g["ex"] // max(g["n"], 1), 100.0 * g["ex"] / tot_ex, 100.0 * g["samp"] / max(tot_s, 1),
        g["thr"] / max(g["ex"], 1), g["atoms"], g["first"][:60]))
