#!/bin/bash
# native training step: tests, timing, then ncu evidence (launch list + a sectioned capture summarised ON the box: the .ncu-rep is too big to bring back)
mkdir -p gpurun_out/r02w
timeout 900 python -m pytest tests/test_gpu_train_native.py tests/test_gpu_parity.py -q -m gpu -k "native or training or impl or stage or variants or accumulate or loop or one_forward" > gpurun_out/r02w/pytest.txt 2>&1
echo "exit $?" >> gpurun_out/r02w/pytest.txt
grep -v "^ \|^$\|^>" gpurun_out/r02w/pytest.txt | tail -12
TRAIN_NO_BRIDGE=1 timeout 300 python tools/train_step_bench.py > gpurun_out/r02w/train_step.txt 2>&1
head -30 gpurun_out/r02w/train_step.txt
K='regex:adam_step|conv1x1_nhwc|conv3x3|convT2x2|conv_final_bwd|conv_in_bwd|conv_in_planes_train|decode_points|pool_bwd|tall_to_nchw|train_pack|xz_finish|giga_loss'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 106 -c 53 --csv --log-file gpurun_out/r02w/train_launches.csv python tools/ncu_train_step.py > gpurun_out/r02w/ncu1.log 2>&1
tail -1 gpurun_out/r02w/ncu1.log
timeout 900 ncu --set full --clock-control none -k "$K" -s 106 -c 53 -o /tmp/train_prof -f python tools/ncu_train_step.py > gpurun_out/r02w/ncu2.log 2>&1
tail -1 gpurun_out/r02w/ncu2.log
python tools/ncu_summary.py /tmp/train_prof.ncu-rep > gpurun_out/r02w/train_step_kernels_ncu_full.txt 2>&1
wc -l gpurun_out/r02w/train_step_kernels_ncu_full.txt
