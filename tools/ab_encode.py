#!/usr/bin/env python
"""Debug: A/B the encoder schedule switches inside ONE process (same box, interleaved repeats): CUDA-event time of
K back-to-back encodes per configuration.  python tools/ab_encode.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, giga_b200
from giga_b200._lib import lib, check
from oracle import giga_oracle as O
B, K = 32, 40
net = giga_b200.get_network("giga"); net.load_state_dict(O.seeded_state_dict(seed=1)); net = net.to("cuda:0")
eng = net._engine()
xs = [torch.rand(B, 40, 40, 40, device="cuda:0") for _ in range(8)]
ps = torch.rand(B, 2048, 3, device="cuda:0") - 0.5
def run(full):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize(); e0.record()
    for i in range(K):
        if full: net(xs[i % 8], ps, p_tsdf=ps)
        else: net.encode_inputs(xs[i % 8])
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K * 1e3
cfgs = [("tile_deps=0", dict(tile_deps=0)), ("tile_deps=1 static", dict(tile_deps=1, dynamic=0)), ("tile_deps=1 dynamic", dict(tile_deps=1, dynamic=1))]
res = {n: [] for n, _ in cfgs}
for rep in range(5):
    for name, c in cfgs:
        eng.set_option("tile_deps", c["tile_deps"])
        if "dynamic" in c: eng.set_option("dynamic_items", c["dynamic"])
        run(False)
        res[name].append((run(False), run(True)))
for name, v in res.items():
    enc = sorted(a for a, _ in v); full = sorted(b for _, b in v)
    print(f"{name:22s} encode {enc[len(enc)//2]:7.1f} us (min {enc[0]:.1f})   forward {full[len(full)//2]:7.1f} us (min {full[0]:.1f})")
