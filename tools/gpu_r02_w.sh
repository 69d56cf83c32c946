#!/bin/bash
# ncu evidence of the native training step: launch list (durations) + full capture of the step's kernels
mkdir -p gpurun_out/r02w
K='regex:giga'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 106 -c 53 --csv --log-file gpurun_out/r02w/train_launches.csv python tools/ncu_train_step.py > gpurun_out/r02w/ncu1.log 2>&1
tail -2 gpurun_out/r02w/ncu1.log
timeout 900 ncu --set full --clock-control none -k "$K" -s 106 -c 53 -o gpurun_out/r02w/train_prof python tools/ncu_train_step.py > gpurun_out/r02w/ncu2.log 2>&1
tail -2 gpurun_out/r02w/ncu2.log
ls -la gpurun_out/r02w/
