#!/bin/bash
mkdir -p gpurun_out/r02z
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r02z/pytest_gpu.txt 2>&1
echo "exit $?" >> gpurun_out/r02z/pytest_gpu.txt
grep -v "^ \|^$\|^>" gpurun_out/r02z/pytest_gpu.txt | tail -15
timeout 300 python tools/train_step_bench.py > gpurun_out/r02z/train_step.txt 2>&1
grep "train step\|launches per" gpurun_out/r02z/train_step.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
