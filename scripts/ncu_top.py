#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: top SASS instructions by stall samples + stall mix.
usage: ncu -i rep --page source --csv > src.csv ; python scripts/ncu_top.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
data = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
S = col["# Samples"]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[S] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
mix = {h: sum(int(r[col[h]] or 0) for r in data) for h in stall_cols}
print("stall mix:", ", ".join(f"{h[6:]}={100*v/max(tot,1):.1f}%" for h, v in sorted(mix.items(), key=lambda kv: -kv[1])[:8]))
order = sorted(range(len(data)), key=lambda i: -int(data[i][S] or 0))[:n]
for i in sorted(order):
    r = data[i]
    top = sorted(((int(r[col[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {100*int(r[S] or 0)/max(tot,1):5.1f}%  {r[col['Source']][:90]:90s} {top}")
