#!/usr/bin/env python
"""One bench step under a profiler: 3 warm-up steps + 1 profiled step of configs[1] (32 scenes, 2048 + 2048 points, all heads + arg-max),
18 kernels per step, nothing else on the GPU.  Used by tools/gpu_r02_evidence.sh:
    ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:conv_in_planes|xz_finish|conv_tall|pool_tall|decode_points|scene_argmax' -s 54 -c 18 --csv --log-file launches.csv python tools/ncu_step.py
    ncu --set full --clock-control none --import-source on -k 'regex:conv_in_planes|xz_finish|conv_tall|pool_tall|decode_points|scene_argmax' -s 54 -c 18 -o prof python tools/ncu_step.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import giga_b200
from oracle import giga_oracle as O

B, N = 32, 2048
dev = torch.device("cuda:0")
net = giga_b200.get_network("giga")
net.load_state_dict(O.seeded_state_dict(seed=1))
net = net.to(dev)
g = torch.Generator(device=dev).manual_seed(1234)
xs = torch.rand((4, B, 40, 40, 40), device=dev, generator=g)
ps = torch.rand((4, B, N, 3), device=dev, generator=g) - 0.5
pts = torch.rand((4, B, N, 3), device=dev, generator=g) - 0.5
val = torch.zeros(B, device=dev)
idx = torch.zeros(B, device=dev, dtype=torch.int32)
with torch.no_grad():
    for i in range(4):
        net.forward_with_argmax(xs[i], ps[i], pts[i], val, idx)
    torch.cuda.synchronize()
print("launches", net.gpu_launches)
