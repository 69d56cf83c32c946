#!/usr/bin/env python
"""A/B: batched planner throughput (giga_detect_host, B = 32) with the parameters committed by the device-side packer vs the host packer."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch, giga_b200
from giga_b200.detection_implicit import detect_host, select_params
from oracle import giga_oracle as O, planner_oracle as PO
tsdfs = np.stack([PO.seeded_volumes(100 + i)[0] for i in range(8)] * 4)
prm = select_params()
for rep in range(2):
    for mode in ("auto", "host"):
        net = giga_b200.get_network("giga"); net = net.to("cuda:0")
        net._engine_raw().commit_mode = mode
        net.load_state_dict(PO.planner_state_dict(O.seeded_state_dict(seed=1)))
        for _ in range(3): detect_host(net, tsdfs, None, prm, K=256)
        t0 = time.perf_counter()
        for _ in range(8): cnt = detect_host(net, tsdfs, None, prm, K=256)[0]
        dt = (time.perf_counter() - t0) / 8
        print(f"commit {mode:4s}: {len(tsdfs) / dt:8.0f} scenes/s ({1e3 * dt:.2f} ms per 32-scene call), grasps {int(cnt.sum())}")
