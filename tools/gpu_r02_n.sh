#!/bin/bash
mkdir -p gpurun_out/r02n
timeout 600 ncu --set full --import-source on --clock-control none -k 'regex:decode_points_bwd|conv_in_bwd' -s 4 -c 2 -o /tmp/p2 -f python tools/ncu_train_step.py > gpurun_out/r02n/ncu.log 2>&1
tail -2 gpurun_out/r02n/ncu.log
python tools/ncu_summary.py /tmp/p2.ncu-rep > gpurun_out/r02n/bwd_kernels_ncu_full.txt 2>&1
cat gpurun_out/r02n/bwd_kernels_ncu_full.txt
ncu -i /tmp/p2.ncu-rep --page source --csv --print-source sass -k regex:decode_points_bwd 2>/dev/null | python - <<'PY' > gpurun_out/r02n/decode_bwd_hot_sass.txt
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr = None
for i, r in enumerate(rows):
    if "Source" in r and any("Sampl" in c for c in r):
        hdr = r; body = rows[i + 1:]; break
if hdr is None:
    print("no source page"); sys.exit()
si = hdr.index("Source")
ci = [j for j, c in enumerate(hdr) if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)"]
ci = ci[0] if ci else None
tot = 0; agg = {}
for r in body:
    try: n = int(r[ci])
    except Exception: continue
    tot += n
    op = r[si].split()[0] if r[si].split() else "?"
    if op.startswith("@"): op = r[si].split()[1]
    agg[op] = agg.get(op, 0) + n
print("total samples", tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:25]:
    print(f"{k:24s} {v:8d} {100.0 * v / max(tot, 1):5.1f}%")
PY
head -30 gpurun_out/r02n/decode_bwd_hot_sass.txt
