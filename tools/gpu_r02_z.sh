#!/bin/bash
mkdir -p gpurun_out/r02z
timeout 900 python -m pytest tests/test_gpu_train_native.py tests/test_gpu_parity.py -q -m gpu > gpurun_out/r02z/pytest.txt 2>&1
echo "exit $?" >> gpurun_out/r02z/pytest.txt
grep -v "^ \|^$\|^>" gpurun_out/r02z/pytest.txt | tail -12
TRAIN_NO_BRIDGE=1 timeout 300 python tools/train_step_bench.py > gpurun_out/r02z/train_step.txt 2>&1
head -40 gpurun_out/r02z/train_step.txt
