#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: stall samples per opcode and the hottest SASS lines.
usage: ncu -i prof.ncu-rep --page source --csv > src.csv; python tools/ncu_hotspots.py src.csv [top]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
src, samp, ex = ci["Source"], ci["Warp Stall Sampling (All Samples)"], ci["Instructions Executed"]
data = [r for r in rows[hi + 1:] if len(r) > samp and r[samp].isdigit()]
tot = sum(int(r[samp]) for r in data)
print(rows[0][1] if len(rows[0]) > 1 else "")
print("total stall samples", tot, "| SASS lines", len(data), "| instructions executed",
      sum(int(r[ex]) for r in data if r[ex].isdigit()))
stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = Counter()
for r in data:
    for h in stall:
        v = r[ci[h]]
        if v.isdigit():
            agg[h] += int(v)
print("stall reasons:", ", ".join(f"{h[6:]} {100 * n / max(tot, 1):.1f}%" for h, n in agg.most_common(9)))
c, e = Counter(), Counter()
for r in data:
    op = r[src].strip().split()
    if not op:
        continue
    o = (op[1] if op[0].startswith("@") else op[0]).split(".")[0].rstrip(";")
    c[o] += int(r[samp])
    e[o] += int(r[ex]) if r[ex].isdigit() else 0
print("\nper opcode: samples (share)  executed")
for o, n in c.most_common(top):
    print(f"  {o:12s} {n:7d} ({100 * n / tot:5.1f}%)  {e[o]}")
print("\nhottest SASS lines:")
for r in sorted(data, key=lambda r: -int(r[samp]))[:top]:
    why = max(stall, key=lambda h: int(r[ci[h]]) if r[ci[h]].isdigit() else 0)
    print(f"  {int(r[samp]):6d}  {why[6:]:14s} {r[src].strip()[:90]}")
