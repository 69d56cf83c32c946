#!/usr/bin/env python
"""BASELINE.json configs[3] (informational): a scripts/train_giga.py-style step (forward + loss + backward + Adam, batch 64, one grasp
point + 2048 occupancy points per sample) through the opt-in training bridge: forward values from the CUDA library, gradients from
the PyTorch recompute on the GPU (giga_b200/training.py).  Prints ms/step and samples/s."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
import giga_b200
from oracle import giga_oracle as O

dev = torch.device("cuda:0")
B, No = 64, 2048
net = giga_b200.get_network("giga"); net.load_state_dict(O.seeded_state_dict(seed=1)); net = net.to(dev).enable_training_bridge()
opt = torch.optim.Adam(net.parameters(), lr=2e-4)
g = torch.Generator(device=dev).manual_seed(0)
x = torch.rand(B, 40, 40, 40, device=dev, generator=g)
pos = torch.rand(B, 1, 3, device=dev, generator=g) - 0.5
pos_occ = torch.rand(B, No, 3, device=dev, generator=g) - 0.5
label = (torch.rand(B, device=dev, generator=g) > 0.5).float()
rot_t = F.normalize(torch.randn(B, 2, 4, device=dev, generator=g), dim=2)
width_t = torch.rand(B, device=dev, generator=g) * 0.1
occ_t = (torch.rand(B, No, device=dev, generator=g) > 0.5).float()

def quat_loss(pred, target):            # train_giga.py:180-182
    return 1.0 - torch.abs(torch.sum(pred * target, dim=1))

def step():
    opt.zero_grad(set_to_none=True)
    qual, rot, width, occ = net(x, pos, p_tsdf=pos_occ)
    qual, rot, width = qual.squeeze(-1), rot.squeeze(1), width.squeeze(-1)
    loss_qual = F.binary_cross_entropy(qual, label, reduction="none")                  # train_giga.py:161-174
    loss_rot = torch.min(quat_loss(rot, rot_t[:, 0]), quat_loss(rot, rot_t[:, 1]))
    loss_width = F.mse_loss(40 * width, 40 * width_t, reduction="none")
    loss_occ = F.binary_cross_entropy(torch.sigmoid(occ), occ_t, reduction="none").mean(-1)
    loss = (loss_qual + label * (loss_rot + 0.01 * loss_width) + loss_occ).mean()
    loss.backward()
    opt.step()
    return loss

for _ in range(3): step()
torch.cuda.synchronize(); t0 = time.perf_counter()
K = 10
for _ in range(K): l = step()
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / K
print(f"train step (bridge) B={B}, 1 grasp pt + {No} occ pts: {1e3 * dt:.2f} ms/step = {B / dt:.0f} samples/s; loss {float(l):.4f}")
with torch.no_grad():
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(K): net(x, pos, p_tsdf=pos_occ)
    torch.cuda.synchronize()
print(f"  forward only (CUDA library): {1e3 * (time.perf_counter() - t0) / K:.2f} ms")
