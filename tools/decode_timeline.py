#!/usr/bin/env python
"""Debug: in-kernel item timeline of the warp-specialised decoder (decoder_ws.cuh).
GIGA_TIMELINE=decode python tools/decode_timeline.py   -> per CTA: stamps of the compute warpgroup (item start, 5 blocks, item end) for its first 4 items"""
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
pt = torch.rand(B, N, 3, device="cuda:0") - 0.5
for _ in range(3):
    out = net(x, p, p_tsdf=pt)       # merged launch: 3 grasp heads + tsdf head
torch.cuda.synchronize()
eng = net._engine()
buf = torch.zeros(8 << 20, dtype=torch.float32, device="cuda:0")
n = lib.giga_debug_copy(eng.h, b"timeline", C.c_void_p(buf.data_ptr()), buf.numel(), C.c_void_p(0))
assert n > 0, lib.giga_last_error()
torch.cuda.synchronize()
t = buf[:n].cpu().numpy().view(np.uint64).reshape(-1, 32).astype(np.int64)
acct = t[:, 21:29].copy()
t[:, 21:32] = 0
t0 = t[t > 0].min()
rel = np.where(t > 0, (t - t0) / 1000.0, np.nan)
print("decode_points_ws CTAs", len(t), "kernel span us %.2f" % np.nanmax(rel))
ends = np.nanmax(rel, axis=1)
print("CTA end times us: min %.1f median %.1f max %.1f" % (ends.min(), np.median(ends), ends.max()))
lab = ["fc_p+blk0", "blk1", "blk2", "blk3", "blk4", "final"]
for k in range(3):
    seg = rel[:, 7 * k:7 * k + 7]
    ok = ~np.isnan(seg).any(axis=1)
    if not ok.any():
        continue
    d = np.diff(seg[ok], axis=1)
    gap = seg[ok, 0] - rel[ok, 7 * k - 1] if k > 0 else seg[ok, 0]
    print(f"item ordinal {k}: {ok.sum()} CTAs, start median {np.median(seg[ok, 0]):.1f} us (gap after previous item {np.median(gap):.2f}), "
          f"duration median {np.median(seg[ok, 6] - seg[ok, 0]):.2f} (min {np.min(seg[ok, 6] - seg[ok, 0]):.2f}, max {np.max(seg[ok, 6] - seg[ok, 0]):.2f})")
    print("   phases: " + "  ".join(f"{l} {np.median(d[:, i]):.2f}" for i, l in enumerate(lab)))
m = np.median(acct, axis=0)
print("MMA-issue lane cycle accounting (median over CTAs), total %.0f cycles:" % m[6])
for n_, v in zip(["issue fc_c batch", "wait a_ready (E1 -> fc_0)", "issue fc_0", "wait a_ready (E2 -> fc_1)", "issue fc_1", "wait item start (features / TMEM free)", None, "wait fc_c weights"], m):
    if n_:
        print("   %-40s %8.0f  (%.1f %%)" % (n_, v, 100 * v / m[6]))
