#!/usr/bin/env python
"""Debug: where the pipelined host path (giga_forward_host_submit/wait) spends its step: PCIe copy rates of the
step's buffers, the submit/wait host times, and the pipelined step time vs the device-only step time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import giga_b200
from oracle import giga_oracle as O

B, N, K = 32, 2048, 60
net = giga_b200.get_network("giga"); net.load_state_dict(O.seeded_state_dict(seed=1)); net = net.to("cuda:0")
pin = lambda *s: torch.empty(s, dtype=torch.float32).pin_memory()
hx = [torch.rand(B, 40, 40, 40).pin_memory() for _ in range(2)]
hp = [(torch.rand(B, N, 3) - 0.5).pin_memory() for _ in range(2)]
hpt = [(torch.rand(B, N, 3) - 0.5).pin_memory() for _ in range(2)]
outs = [(pin(B, N), pin(B, N, 4), pin(B, N), pin(B, N)) for _ in range(2)]
xd = torch.empty(B, 40, 40, 40, device="cuda")
od = torch.empty(B, N, 7, device="cuda"); oh = pin(B, N, 7)
for name, fn, nbytes in (("H2D tsdf 8.2MB", lambda: xd.copy_(hx[0], non_blocking=True), xd.numel() * 4),
                         ("D2H out 1.8MB", lambda: oh.copy_(od, non_blocking=True), od.numel() * 4)):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{name}: {ms:.3f} ms  {nbytes / ms / 1e6:.1f} GB/s")
xdv, pdv, ptdv = hx[0].cuda(), hp[0].cuda(), hpt[0].cuda()
for _ in range(5): net(xdv, pdv, p_tsdf=ptdv)
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(K): net(xdv, pdv, p_tsdf=ptdv)
torch.cuda.synchronize(); print(f"device-only step {1e3 * (time.perf_counter() - t0) / K:.3f} ms")
for i in range(4):
    net.forward_host_submit(i % 2, hx[i % 2], hp[i % 2], hpt[i % 2], outs[i % 2]); net.forward_host_wait(i % 2)
for rep in range(3):
    sub, wai = [], []
    t0 = time.perf_counter()
    for i in range(K):
        if i >= 2:
            a = time.perf_counter(); net.forward_host_wait(i % 2); wai.append(time.perf_counter() - a)
        a = time.perf_counter(); net.forward_host_submit(i % 2, hx[i % 2], hp[i % 2], hpt[i % 2], outs[i % 2]); sub.append(time.perf_counter() - a)
    for i in range(K - 2, K): net.forward_host_wait(i % 2)
    pipe_ms = 1e3 * (time.perf_counter() - t0) / K
    med = lambda v: 1e3 * sorted(v)[len(v) // 2]
    print(f"pipelined {pipe_ms:.3f} ms/step | submit median {med(sub):.3f} max {1e3 * max(sub):.3f} | wait median {med(wai):.3f}")
t0 = time.perf_counter()
for i in range(20): net.forward_host(hx[0], hp[0], hpt[0], out=outs[0])
print(f"sync host call {1e3 * (time.perf_counter() - t0) / 20:.3f} ms")
