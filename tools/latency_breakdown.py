#!/usr/bin/env python
"""Debug: per-kernel CUDA-event times of one call at small batch (the planner's B=1 / 64,000-point forward + post-processing,
and the C1 configuration B=1 / 2048+2048 points), plus the un-instrumented wall-clock latency of the same call."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
import giga_b200
from giga_b200.detection_implicit import detect_host, select_params
from oracle import giga_oracle as O, planner_oracle as P

dev = torch.device("cuda:0")
net = giga_b200.get_network("giga"); net.load_state_dict(P.planner_state_dict(O.seeded_state_dict(seed=1))); net = net.to(dev)
eng = net._engine()
tsdf = P.seeded_volumes(5)[0][None]
x, p, pt = (t.to(dev) for t in O.seeded_inputs(1, 2048, seed=0))
jobs = {"detect B=1 (64000-pt lattice, 3 heads, post-processing)": lambda: detect_host(net, tsdf, None, select_params(), K=256),
        "forward B=1 (2048+2048 pts, 4 heads)": lambda: (net(x, p, p_tsdf=pt), torch.cuda.synchronize())}
for name, fn in jobs.items():
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(50): fn()
    torch.cuda.synchronize(); wall = 1e3 * (time.perf_counter() - t0) / 50
    eng.set_timing(True)
    for _ in range(10): fn()
    torch.cuda.synchronize()
    eng.set_timing(False)
    rep = eng.timing_report()
    tot = sum(ms for _, ms in rep.values()) / 10
    print(f"{name}: wall {wall:.3f} ms/call; sum of kernel times {tot:.3f} ms ({len(rep)} distinct kernels)")
    for k, (n, ms) in rep.items():
        print(f"    {k:28s} {1e3 * ms / n:8.1f} us x{n // 10}")
