#!/usr/bin/env python
"""Top stall-sample SASS instructions of one kernel in an .ncu-rep (source page).
usage: python tools/ncu_hot.py <rep> <kernel-regex> [launch-index] [topN]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
# the output is a sequence of tables, one per kernel launch, each starting with a "Kernel Name" row
tables, cur = [], None
for row in csv.reader(io.StringIO(raw)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}
        tables.append(cur)
    elif cur is not None:
        cur["rows"].append(row)
t = tables[which]
hdr = t["rows"][0]
col = {h: i for i, h in enumerate(hdr)}
rows = [r for r in t["rows"][1:] if len(r) == len(hdr)]
tot = sum(int(r[col["# Samples"]] or 0) for r in rows)
print(t["name"][:120], "total samples", tot)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
best = sorted(rows, key=lambda r: -int(r[col["# Samples"]] or 0))[:topn]
idx = {id(r): i for i, r in enumerate(rows)}
for r in best:
    n = int(r[col["# Samples"]] or 0)
    st = sorted(((int(r[col[s]] or 0), s[6:]) for s in stall_cols), reverse=True)[:2]
    print(f"{n:6d} {100.0*n/max(tot,1):5.1f}%  #{idx[id(r)]:5d}  {r[col['Source']].strip()[:90]:<90} {st}")
