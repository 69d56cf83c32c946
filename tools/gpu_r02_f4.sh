#!/bin/bash
# round-2 f4 validation: VGN baseline + loss/Adam kernels, then the whole GPU suite
mkdir -p gpurun_out/r02r
timeout 900 python -m pytest tests/test_gpu_vgn.py tests/test_gpu_train.py -x -q -m gpu > gpurun_out/r02r/pytest_f4.txt 2>&1
echo "f4 exit $?" >> gpurun_out/r02r/pytest_f4.txt
tail -30 gpurun_out/r02r/pytest_f4.txt
timeout 300 python - > gpurun_out/r02r/vgn_timing.txt 2>&1 <<'PY'
import torch, giga_b200
from oracle import vgn_oracle as V
net = giga_b200.get_network("vgn"); net.load_state_dict(V.seeded_state_dict(3)); net = net.cuda()
for B in (1, 32):
    x = V.seeded_inputs(min(B, 4), 1).cuda().repeat((B + 3) // 4, 1, 1, 1, 1)[:B].contiguous()
    net._engine().set_timing(False)
    for _ in range(3): net.forward_flat(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): net.forward_flat(x)
    e1.record(); torch.cuda.synchronize()
    print(f"B={B}: {e0.elapsed_time(e1) / 10:.3f} ms per forward")
    net._engine().set_timing(True)
    net.forward_flat(x); torch.cuda.synchronize()
    for k, v in net._engine().timing_report().items(): print("   ", k, v)
    net._engine().set_timing(False)
PY
cat gpurun_out/r02r/vgn_timing.txt
