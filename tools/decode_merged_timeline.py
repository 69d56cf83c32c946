#!/usr/bin/env python
"""Debug: CTA-level timeline of the merged decoder launch (grasp heads + TSDF head in one grid, cost-ordered).
GIGA_TIMELINE=decode python tools/decode_merged_timeline.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch, giga_b200
from giga_b200._lib import lib
from oracle import giga_oracle as O
B, N = 32, 2048
net = giga_b200.get_network("giga"); net.load_state_dict(O.seeded_state_dict(seed=1)); net = net.to("cuda:0")
x = torch.rand(B, 40, 40, 40, device="cuda:0"); p = torch.rand(B, N, 3, device="cuda:0") - 0.5; pt = torch.rand(B, N, 3, device="cuda:0") - 0.5
for _ in range(3): net(x, p, p_tsdf=pt)
torch.cuda.synchronize()
buf = torch.zeros(8 << 20, dtype=torch.float32, device="cuda:0")
n = lib.giga_debug_copy(net._engine().h, b"timeline", C.c_void_p(buf.data_ptr()), buf.numel(), C.c_void_p(0))
assert n > 0, lib.giga_last_error()
torch.cuda.synchronize()
t = buf[:n].cpu().numpy().view(np.uint64).reshape(-1, 32).astype(np.int64)
start = t[:, 0]
end = np.where(t[:, :31] > 0, t[:, :31], 0).max(1)        # last stamp of the CTA
t0 = start.min()
s, e = (start - t0) / 1e3, (end - t0) / 1e3
nh = 16 * B
print(f"CTAs {len(t)} (heavy {nh}, light {len(t) - nh}); span {e.max():.1f} us")
print(f"heavy: start median {np.median(s[:nh]):.1f} max {s[:nh].max():.1f}; duration median {np.median(e[:nh] - s[:nh]):.1f}; last end {e[:nh].max():.1f}")
print(f"light: start min {s[nh:].min():.1f} median {np.median(s[nh:]):.1f} max {s[nh:].max():.1f}; duration median {np.median(e[nh:] - s[nh:]):.1f}; last end {e[nh:].max():.1f}")
for lo in range(0, int(e.max()) + 10, 10):
    act = ((s <= lo) & (e > lo))
    print(f"  t={lo:4d} us: active heavy {int(act[:nh].sum()):4d}  light {int(act[nh:].sum()):4d}")
