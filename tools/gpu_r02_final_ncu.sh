#!/bin/bash
# final evidence of the inference step: ncu launch list + full capture of the 18 kernels, summarised ON the box (the .ncu-rep stays there)
OUT=gpurun_out/r02i
mkdir -p $OUT
K='regex:conv_in_planes|xz_finish|conv_tall|pool_tall|decode_points|scene_argmax'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 54 -c 18 --csv --log-file $OUT/launches.csv python tools/ncu_step.py > $OUT/ncu_launch.log 2>&1
tail -1 $OUT/ncu_launch.log
timeout 1200 ncu --set full --clock-control none -k "$K" -s 54 -c 18 -o /tmp/prof -f python tools/ncu_step.py > $OUT/ncu_full.log 2>&1
tail -1 $OUT/ncu_full.log
python tools/ncu_summary.py /tmp/prof.ncu-rep > $OUT/step_kernels_ncu_full.txt 2>&1
cp profiles/traffic.json /tmp/traffic_old.json
python tools/ncu_traffic.py /tmp/prof.ncu-rep > $OUT/traffic.log 2>&1
cp profiles/traffic.json $OUT/traffic.json
python - <<PY > $OUT/launch_shares.txt
import csv
rows = [r for r in csv.reader(open("gpurun_out/r02i/launches.csv")) if len(r) > 10 and r[0].isdigit()]
tot = sum(float(r[-1]) for r in rows)
print("ncu launch list of one step (gpu__time_duration.sum, ns; serialised, cold caches): total %.1f us, %d launches" % (tot / 1e3, len(rows)))
for r in rows: print("%-70s %9.1f us %5.1f %%" % (r[4][:70], float(r[-1]) / 1e3, 100 * float(r[-1]) / tot))
PY
wc -l $OUT/*.txt $OUT/launches.csv
