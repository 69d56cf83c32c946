#!/bin/bash
# usage (under gpurun): bash tools/gpu_ncu.sh <tag> <kernel-regex> <skip> <count>
TAG=$1; RX=$2; SKIP=${3:-0}; CNT=${4:-10}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RX" -s $SKIP -c $CNT -o $OUT/prof -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu.log 2>&1
tail -3 $OUT/ncu.log
ls -la $OUT
