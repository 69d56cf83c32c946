#!/usr/bin/env python
"""Time the BASELINE.json configurations (forward only) through the public API on one GPU, CUDA events,
median of R runs after warm-up, and (optionally) the CPU oracle on the same inputs.
usage: python tools/config_sweep.py [--cpu]"""
import json, os, statistics, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import giga_b200
from oracle import giga_oracle as O

dev = torch.device("cuda:0")
sd = O.seeded_state_dict(seed=1)
net = giga_b200.get_network("giga"); net.load_state_dict(sd); net = net.to(dev)
lin = torch.linspace(-0.5, 0.5 - 1.0 / 40, 40)
lattice = torch.stack(torch.meshgrid(lin, lin, lin, indexing="ij"), -1).reshape(1, -1, 3)   # the planner's 40^3 lattice
CFG = [
    ("C1  B=1, 2048+2048 pts, 4 heads", 1, 2048, 2048, "full"),
    ("C2  B=32, 2048+2048 pts, 4 heads", 32, 2048, 2048, "full"),
    ("C3  B=32 (per GPU), 4096+4096 pts, 4 heads", 32, 4096, 4096, "full"),
    ("C5  B=1, 64000 pts, occupancy head only (infer_geo)", 1, 0, 64000, "geo"),
    ("C5' B=8, 64000 pts, occupancy head only", 8, 0, 64000, "geo"),
    ("sim B=1, 64000-pt lattice, 3 grasp heads (VGNImplicit)", 1, 64000, 0, "grasp"),
]
rows = []
for name, B, Ng, No, mode in CFG:
    g = torch.Generator().manual_seed(0)
    x = torch.rand(B, 40, 40, 40, generator=g)
    p = (lattice.expand(B, -1, -1).contiguous() if Ng == 64000 else torch.rand(B, max(Ng, 1), 3, generator=g) - 0.5)
    pt = (lattice.expand(B, -1, -1).contiguous() if No == 64000 else torch.rand(B, max(No, 1), 3, generator=g) - 0.5)
    xd, pd, ptd = x.to(dev), p.to(dev), pt.to(dev)
    fn = {"full": lambda: net(xd, pd, p_tsdf=ptd), "geo": lambda: net.infer_geo(xd, ptd), "grasp": lambda: net(xd, pd)}[mode]
    with torch.no_grad():
        for _ in range(5): fn()
        ts = []
        for _ in range(30):
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = statistics.median(ts)
    pts = B * (Ng + No)
    row = {"config": name, "ms": round(ms, 4), "scenes_per_s": round(B / ms * 1e3, 1), "points_per_s": round(pts / ms * 1e3)}
    if "--cpu" in sys.argv:
        torch.set_num_threads(min(32, os.cpu_count()))
        cf = {"full": lambda: O.forward(sd, x, p, pt), "geo": lambda: O.infer_geo(sd, x, pt), "grasp": lambda: O.forward(sd, x, p)}[mode]
        with torch.no_grad():
            cf(); t0 = time.perf_counter(); cf(); cms = 1e3 * (time.perf_counter() - t0)
        row["cpu_oracle_ms"] = round(cms, 1); row["speedup"] = round(cms / ms, 1)
    rows.append(row); print(json.dumps(row), flush=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "config_sweep.json"), "w"), indent=1)
