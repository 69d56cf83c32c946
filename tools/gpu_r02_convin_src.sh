#!/bin/bash
OUT=gpurun_out/r02j
mkdir -p $OUT
timeout 600 ncu --set full --import-source on --clock-control none -k 'regex:conv_in_planes' -s 3 -c 1 -o /tmp/ci -f python tools/ncu_step.py > $OUT/ncu.log 2>&1
tail -1 $OUT/ncu.log
ncu -i /tmp/ci.ncu-rep --page source --csv --print-source sass > /tmp/ci_src.csv 2>/dev/null
python - <<'PY' > gpurun_out/r02j/conv_in_smem_conflicts.txt
import csv
rows = list(csv.reader(open("/tmp/ci_src.csv")))
hdr = next(r for r in rows if "Source" in r)
i0 = rows.index(hdr)
col = {h: i for i, h in enumerate(hdr)}
print([h for h in hdr if "onflict" in h or "avefront" in h or "Sampl" in h])
keys = [h for h in hdr if ("Shared" in h or "shared" in h) and ("onflict" in h or "avefront" in h or "Excessive" in h)]
out = []
for r in rows[i0 + 1:]:
    if len(r) != len(hdr): continue
    vals = []
    for k in keys:
        try: vals.append(float(r[col[k]] or 0))
        except ValueError: vals.append(0.0)
    if any(v > 0 for v in vals): out.append((vals, r[col["Source"]].strip()[:100]))
print(keys)
out.sort(key=lambda t: -max(t[0]))
for vals, src in out[:40]: print(["%.0f" % v for v in vals], src)
PY
head -50 gpurun_out/r02j/conv_in_smem_conflicts.txt
