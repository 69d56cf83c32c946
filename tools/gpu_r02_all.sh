#!/bin/bash
# whole GPU suite + the native training step timing
mkdir -p gpurun_out/r02t
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r02t/pytest_gpu.txt 2>&1
echo "pytest-gpu exit $?" >> gpurun_out/r02t/pytest_gpu.txt
grep -v "^ \|^$\|^E  \|^>" gpurun_out/r02t/pytest_gpu.txt | tail -30
TRAIN_NO_BRIDGE=1 timeout 300 python tools/train_step_bench.py > gpurun_out/r02t/train_step.txt 2>&1
head -14 gpurun_out/r02t/train_step.txt
