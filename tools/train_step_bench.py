#!/usr/bin/env python
"""BASELINE.json configs[3]: a scripts/train_giga.py-style step (forward + loss + backward + Adam, batch 64, one grasp point + 2048
occupancy points per sample) on ONE GPU.  Three arms, same data:
  native   giga_train_forward / giga_train_backward (csrc/train_bwd.cuh) + fused loss + flat Adam (csrc/train.cuh): this library only
  bridge   round 1's path, now test infrastructure (tests/torch_bridge.py: forward by the library, backward = PyTorch recompute on ATen/cuDNN) + torch ops for loss/Adam
Prints ms/step, samples/s and the per-kernel device times of one native step."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
import giga_b200
from giga_b200 import training
from oracle import giga_oracle as O

dev = torch.device("cuda:0")
B, No = int(os.environ.get("TRAIN_B", 64)), 2048
g = torch.Generator(device=dev).manual_seed(0)
x = torch.rand(B, 40, 40, 40, device=dev, generator=g)
pos = torch.rand(B, 1, 3, device=dev, generator=g) - 0.5
pos_occ = torch.rand(B, No, 3, device=dev, generator=g) - 0.5
label = (torch.rand(B, device=dev, generator=g) > 0.5).float()
rot_t = F.normalize(torch.randn(B, 2, 4, device=dev, generator=g), dim=2)
width_t = torch.rand(B, device=dev, generator=g) * 0.1
occ_t = (torch.rand(B, No, device=dev, generator=g) > 0.5).float()
y = (label, rot_t, width_t, occ_t)


def make(bridge):
    net = giga_b200.get_network("giga"); net.load_state_dict(O.seeded_state_dict(seed=1)); net = net.to(dev)
    return net


def timed(fn, K=10, W=3):
    for _ in range(W): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(K): out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K, 1e3 * (time.perf_counter() - t0) / K, out


res = {"B": B, "No": No}
# ---- native ----
net = make(False)
opt = training.Adam(net.parameters(), lr=2e-4)
def native_step():
    opt.zero_grad()
    loss, _ = training.loss_fn(training.select(net(x, pos, p_tsdf=pos_occ)), y)
    loss.backward()
    opt.step()
    return loss
dev_ms, wall_ms, l = timed(native_step)
res["native"] = {"ms_per_step_device": dev_ms, "ms_per_step_wall": wall_ms, "samples_per_s": B / (wall_ms / 1e3), "loss": float(l)}
print(f"native train step B={B}: {dev_ms:.2f} ms (device), {wall_ms:.2f} ms (wall) = {B / wall_ms * 1e3:.0f} samples/s; loss {float(l):.4f}")
eng = net._engine_raw()
n0 = eng.launches
native_step(); torch.cuda.synchronize()
res["native"]["launches_per_step"] = eng.launches - n0
eng.set_timing(True)
native_step(); torch.cuda.synchronize()
rep = eng.timing_report(); eng.set_timing(False)
res["native"]["kernels_us"] = {k: round(1e3 * v[1], 1) for k, v in rep.items()}
tot = sum(v[1] for v in rep.values())
print(f"  launches per step: {res['native']['launches_per_step']}; kernel time by CUDA events (serialised): {tot:.2f} ms")
for k, v in sorted(rep.items(), key=lambda kv: -kv[1][1]):
    print(f"    {k:28s} x{v[0]:<2d} {1e3 * v[1]:9.1f} us")
def fwd_only():
    with torch.no_grad():
        return net(x, pos, p_tsdf=pos_occ)
d, w, _ = timed(fwd_only)
print(f"  inference forward (tcgen05 path, re-commit once): {d:.2f} ms")
del net, opt
# ---- bridge (round-1 path) ----
if not os.environ.get("TRAIN_NO_BRIDGE"):
    from tests.torch_bridge import bridged_forward
    net = make(True)
    topt = torch.optim.Adam(net.parameters(), lr=2e-4)
    def bridge_step():
        topt.zero_grad(set_to_none=True)
        qual, rot, width, occ = bridged_forward(net, x, pos, pos_occ)
        qual, rot, width = qual.squeeze(-1), rot.squeeze(1), width.squeeze(-1)
        l_qual = F.binary_cross_entropy(qual, label, reduction="none")
        l_rot = torch.min(1.0 - (rot * rot_t[:, 0]).sum(1).abs(), 1.0 - (rot * rot_t[:, 1]).sum(1).abs())
        l_width = F.mse_loss(40 * width, 40 * width_t, reduction="none")
        l_occ = F.binary_cross_entropy(torch.sigmoid(occ), occ_t, reduction="none").mean(-1)
        loss = (l_qual + label * (l_rot + 0.01 * l_width) + l_occ).mean()
        loss.backward()
        topt.step()
        return loss
    dev_ms, wall_ms, l = timed(bridge_step, K=5, W=2)
    res["bridge"] = {"ms_per_step_device": dev_ms, "ms_per_step_wall": wall_ms, "samples_per_s": B / (wall_ms / 1e3), "loss": float(l)}
    print(f"bridge train step B={B}: {dev_ms:.2f} ms (device), {wall_ms:.2f} ms (wall) = {B / wall_ms * 1e3:.0f} samples/s; loss {float(l):.4f}")
print(json.dumps(res))
