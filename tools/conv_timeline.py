#!/usr/bin/env python
"""Debug: in-kernel phase timeline of one tensor-core conv layer (run on the GPU box).
usage: GIGA_TIMELINE='conv3x3:d0c1' python tools/conv_timeline.py [B]"""
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

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
net = giga_b200.get_network("giga")
net.load_state_dict(O.seeded_state_dict(seed=1))
net = net.to("cuda:0")
x = torch.rand(B, 40, 40, 40, device="cuda:0")
for _ in range(3):
    net.encode_inputs(x)
torch.cuda.synchronize()
eng = net._engine()
buf = torch.zeros(8 << 20, dtype=torch.float32, device="cuda:0")
n = lib.giga_debug_copy(eng.h, b"timeline", C.c_void_p(buf.data_ptr()), buf.numel(), C.c_void_p(0))
assert n > 0, lib.giga_last_error()
torch.cuda.synchronize()
t = buf[:n].cpu().numpy().view(np.uint64).reshape(-1, 32).astype(np.int64)
t0 = t[:, 0].min()
rel = (t - t0) / 1000.0  # us
names = {0: "start", 1: "ctl:setup", 2: "full0", 3: "mma0", 4: "full1", 5: "mma1", 6: "full2", 7: "mma2", 8: "full3", 9: "mma3", 10: "ctl:done",
         12: "acc0", 13: "acc1", 14: "acc2", 15: "acc3", 20: "drained", 21: "epi:done", 22: "end"}
print(os.environ.get("GIGA_TIMELINE"), "CTAs", len(t), "kernel span us", rel[:, 22].max())
order = np.argsort(t[:, 0])
for label, idx in (("first CTA", order[0]), ("median CTA", order[len(order) // 2]), ("last CTA", order[-1])):
    r = rel[idx]
    print(f"-- {label} (cta {idx}, sm {t[idx,31]}):", " ".join(f"{names[k]}={r[k]-r[0]:.2f}" for k in sorted(names) if t[idx, k] > 0))
dur = rel[:, 22] - rel[:, 0]
print("CTA duration us: min %.2f median %.2f max %.2f" % (dur.min(), np.median(dur), dur.max()))
for k in (2, 4, 6, 8, 10, 20, 21):
    d = rel[:, k] - rel[:, 0]
    print(f"  {names[k]:>10}: median {np.median(d):.2f}  p90 {np.percentile(d, 90):.2f}")
starts = np.sort(rel[:, 0])
print("CTA start times us (every 50th):", np.round(starts[::50], 1))
