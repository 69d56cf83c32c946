#!/usr/bin/env python
"""Debug: host-path (pipelined submit/wait) step time vs encoder implementation, with per-call host timings."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import giga_b200
from oracle import giga_oracle as O

B, N, K = 32, 2048, 30
net = giga_b200.get_network("giga"); net.load_state_dict(O.seeded_state_dict(seed=1)); net = net.to("cuda:0")
pin = lambda *s: torch.empty(s, dtype=torch.float32).pin_memory()
hx = [torch.rand(B, 40, 40, 40).pin_memory() for _ in range(2)]
hp = [(torch.rand(B, N, 3) - 0.5).pin_memory() for _ in range(2)]
hpt = [(torch.rand(B, N, 3) - 0.5).pin_memory() for _ in range(2)]
outs = [(pin(B, N), pin(B, N, 4), pin(B, N), pin(B, N)) for _ in range(2)]
for impl in (1, 0):
    net._engine().set_option("encoder_impl", impl)
    for i in range(4):
        net.forward_host(hx[i % 2], hp[i % 2], hpt[i % 2], out=outs[0])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K):
        net.forward_host(hx[i % 2], hp[i % 2], hpt[i % 2], out=outs[0])
    sync_ms = 1e3 * (time.perf_counter() - t0) / K
    # pipelined
    for i in range(2):
        net.forward_host_submit(i, hx[i], hp[i], hpt[i], outs[i])
    for i in range(2):
        net.forward_host_wait(i)
    sub, wai = [], []
    t0 = time.perf_counter()
    for i in range(K):
        if i >= 2:
            a = time.perf_counter(); net.forward_host_wait(i % 2); wai.append(time.perf_counter() - a)
        a = time.perf_counter(); net.forward_host_submit(i % 2, hx[i % 2], hp[i % 2], hpt[i % 2], outs[i % 2]); sub.append(time.perf_counter() - a)
    for i in range(K - 2, K):
        net.forward_host_wait(i % 2)
    pipe_ms = 1e3 * (time.perf_counter() - t0) / K
    # device-only
    xd, pd, ptd = hx[0].cuda(), hp[0].cuda(), hpt[0].cuda()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(K):
        net(xd, pd, p_tsdf=ptd)
    torch.cuda.synchronize(); dev_ms = 1e3 * (time.perf_counter() - t0) / K
    print(f"encoder_impl={impl}: device {dev_ms:.3f} ms  sync-host {sync_ms:.3f} ms  pipelined {pipe_ms:.3f} ms | submit cpu ms median {1e3*sorted(sub)[len(sub)//2]:.3f} max {1e3*max(sub):.3f} | wait ms median {1e3*sorted(wai)[len(wai)//2]:.3f}")
