#!/usr/bin/env python
"""Debug: in-kernel phase timeline of the tensor-core decoder.  GIGA_TIMELINE=decode python tools/decode_timeline.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import giga_b200
from giga_b200._lib import lib
from oracle import giga_oracle as O

B, N = 32, 2048
net = giga_b200.get_network("giga")
net.load_state_dict(O.seeded_state_dict(seed=1))
net = net.to("cuda:0")
x = torch.rand(B, 40, 40, 40, device="cuda:0")
p = torch.rand(B, N, 3, device="cuda:0") - 0.5
c = net.encode_inputs(x)
for _ in range(3):
    out = net.decode(p, c)       # 3 grasp heads
torch.cuda.synchronize()
eng = net._engine()
buf = torch.zeros(8 << 20, dtype=torch.float32, device="cuda:0")
n = lib.giga_debug_copy(eng.h, b"timeline", C.c_void_p(buf.data_ptr()), buf.numel(), C.c_void_p(0))
assert n > 0, lib.giga_last_error()
torch.cuda.synchronize()
t = buf[:n].cpu().numpy().view(np.uint64).reshape(-1, 32).astype(np.int64)
t0 = t[:, 0].min()
rel = (t - t0) / 1000.0
order = np.argsort(t[:, 0])
print("decode (3 heads) CTAs", len(t), "span us", rel[:, :26].max())
for label, idx in (("first", order[0]), ("median", order[len(order) // 2]), ("last", order[-1])):
    r = rel[idx] - rel[idx, 0]
    k = [i for i in range(26) if t[idx, i] > 0]
    print(f"-- {label} CTA {idx}: " + " ".join(f"{r[i]:.2f}" for i in k))
d = rel[:, 1:26] - rel[:, 0:25]
lab = ["gather"] + sum([[f"h{h}.pl{p}" for p in range(3)] + [f"h{h}.blk{b}" for b in range(5)] for h in range(3)], [])
print("median phase durations us:")
for i, l in enumerate(lab):
    print(f"  {l:>8}: {np.median(d[:, i]):.2f}", end="" if (i + 1) % 4 else "\n")
print()
print("CTA start times (every 64th):", np.round(np.sort(rel[:, 0])[::64], 1))
