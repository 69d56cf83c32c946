#!/bin/bash
mkdir -p gpurun_out/r02x
timeout 900 python -m pytest tests/test_gpu_train_native.py -q -m gpu > gpurun_out/r02x/pytest.txt 2>&1
echo "exit $?" >> gpurun_out/r02x/pytest.txt
grep -v "^ \|^$\|^>" gpurun_out/r02x/pytest.txt | tail -8
TRAIN_NO_BRIDGE=1 timeout 300 python tools/train_step_bench.py > gpurun_out/r02x/train_step.txt 2>&1
head -14 gpurun_out/r02x/train_step.txt
