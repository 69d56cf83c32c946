#!/bin/bash
mkdir -p gpurun_out/r02y
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r02y/pytest_gpu.txt 2>&1
echo "exit $?" >> gpurun_out/r02y/pytest_gpu.txt
grep -v "^ \|^$\|^>" gpurun_out/r02y/pytest_gpu.txt | tail -25
python tools/commit_time.py > gpurun_out/r02y/commit_time.txt 2>&1; tail -5 gpurun_out/r02y/commit_time.txt
