#!/bin/bash
mkdir -p gpurun_out/r02q
timeout 900 python -m pytest tests/test_gpu_train_native.py -q -m gpu > gpurun_out/r02q/pytest.txt 2>&1
echo "exit $?" >> gpurun_out/r02q/pytest.txt
grep -v "^ \|^$\|^>" gpurun_out/r02q/pytest.txt | tail -6
TRAIN_NO_BRIDGE=1 timeout 300 python tools/train_step_bench.py > gpurun_out/r02q/train_step.txt 2>&1
grep "train step\|launches per\|conv_in_bwd\|decode_bwd" gpurun_out/r02q/train_step.txt | head
