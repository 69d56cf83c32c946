#!/usr/bin/env python
"""One native training step (configs[3]: B = 64, 1 grasp point + 2048 occupancy points per sample) under a profiler: 2 warm-up steps +
1 profiled step (53 kernels: pack, forward, loss, backward, Adam).
    ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:<the step kernels, see tools/gpu_r02_w.sh>' -s 106 -c 53 --csv --log-file launches.csv python tools/ncu_train_step.py
    ncu --set full --clock-control none -k 'regex:<the step kernels, see tools/gpu_r02_w.sh>' -s 106 -c 53 -o prof python tools/ncu_train_step.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
import giga_b200
from giga_b200 import training
from oracle import giga_oracle as O

dev = torch.device("cuda:0")
B, No = 64, 2048
g = torch.Generator(device=dev).manual_seed(0)
x = torch.rand(B, 40, 40, 40, device=dev, generator=g)
pos = torch.rand(B, 1, 3, device=dev, generator=g) - 0.5
pos_occ = torch.rand(B, No, 3, device=dev, generator=g) - 0.5
y = ((torch.rand(B, device=dev, generator=g) > 0.5).float(), F.normalize(torch.randn(B, 2, 4, device=dev, generator=g), dim=2),
     torch.rand(B, device=dev, generator=g) * 0.1, (torch.rand(B, No, device=dev, generator=g) > 0.5).float())
net = giga_b200.get_network("giga"); net.load_state_dict(O.seeded_state_dict(seed=1)); net = net.to(dev)
opt = training.Adam(net.parameters(), lr=2e-4)
for _ in range(3):
    opt.zero_grad()
    loss, _ = training.loss_fn(training.select(net(x, pos, p_tsdf=pos_occ)), y)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print("launches", net._engine_raw().launches, "loss", float(loss.detach()))
