#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu launch list + full captures of the top kernels.
# usage (under gpurun):  bash tools/gpu_round.sh <tag> [skip_ncu]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
nproc > $OUT/nproc.txt
echo "== pytest -m gpu" | tee $OUT/pytest_gpu.txt
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee -a $OUT/pytest_gpu.txt
echo "== smoke" | tee $OUT/smoke.txt
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee -a $OUT/smoke.txt
echo "== bench"
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?" >> $OUT/bench.err
tail -c 3000 $OUT/bench.json; tail -3 $OUT/bench.err
if [ -z "$2" ]; then
  echo "== ncu launch list"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file $OUT/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch.log 2>&1
  echo "== ncu full: U-Net kernels of one step (tcgen05 convs, pools, layout)"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_tall_persistent|pool_tall|nchw_to_tall' -s 45 -c 15 -o $OUT/prof_unet -f \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_unet.log 2>&1
  echo "== ncu full: conv_in + decoder + argmax of one step"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'decode_points|conv_in_planes|xz_finish|scene_argmax' -s 15 -c 5 -o $OUT/prof_misc -f \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_misc.log 2>&1
  ls -la $OUT
fi
