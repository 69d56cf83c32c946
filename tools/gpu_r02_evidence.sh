#!/bin/bash
# Round-2 evidence run (under gpurun): launch list + ncu --set full of every kernel of one step, conv_in channel-split A/B.
TAG=${1:-r02p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ -n "$AB" ]; then
echo "== conv_in split A/B"
for s in 1 2; do
  GIGA_CONV_IN_SPLIT=$s timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $OUT/bench_split$s.json 2> $OUT/bench_split$s.err
  python - <<PY
import json
d = json.loads(open("$OUT/bench_split$s.json").read().strip().splitlines()[-1])
print("split $s: ms/step", d["ms_per_step"], "conv_in", d["kernels"]["conv_in_planes"]["us"], "decoder", d["kernels"]["decode_points:grasp+tsdf"]["us"])
PY
done
fi
echo "== ncu launch list (one step)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:conv_in_planes|xz_finish|conv_tall|pool_tall|decode_points|scene_argmax' -s 54 -c 18 --csv --log-file $OUT/launches.csv python tools/ncu_step.py > $OUT/ncu_launch.log 2>&1
tail -2 $OUT/ncu_launch.log
echo "== ncu --set full (one step, 18 kernels)"
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:conv_in_planes|xz_finish|conv_tall|pool_tall|decode_points|scene_argmax' -s 54 -c 18 -o $OUT/prof -f python tools/ncu_step.py > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log
ls -la $OUT
