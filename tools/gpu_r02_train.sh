#!/bin/bash
# round-2 native training step: parity tests (gradients vs CPU autograd through the oracle), then a timed train_giga.py-style step
mkdir -p gpurun_out/r02s
B=9 NO=300 BRIDGE=1 timeout 300 python tools/train_grad_report.py > gpurun_out/r02s/grad_report_b9.txt 2>&1
grep "BAD\|forward\|Error\|bridge" gpurun_out/r02s/grad_report_b9.txt | head -50
timeout 900 python -m pytest tests/test_gpu_train_native.py -q -m gpu > gpurun_out/r02s/pytest_train_native.txt 2>&1
echo "train-native exit $?" >> gpurun_out/r02s/pytest_train_native.txt
grep -v "^ \|^$\|^E  \|^>" gpurun_out/r02s/pytest_train_native.txt | tail -25
TRAIN_NO_BRIDGE=1 timeout 300 python tools/train_step_bench.py > gpurun_out/r02s/train_step.txt 2>&1
head -12 gpurun_out/r02s/train_step.txt
