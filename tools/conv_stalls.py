#!/usr/bin/env python
"""Debug: stall accounting of one persistent conv layer.  GIGA_TIMELINE='conv3x3:d0c1' python tools/conv_stalls.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch, giga_b200
from giga_b200._lib import lib
from oracle import giga_oracle as O
B = 32
net = giga_b200.get_network("giga"); net.load_state_dict(O.seeded_state_dict(seed=1)); net = net.to("cuda:0")
x = torch.rand(B, 40, 40, 40, device="cuda:0")
for _ in range(3): net.encode_inputs(x)
torch.cuda.synchronize()
buf = torch.zeros(1 << 20, dtype=torch.float32, device="cuda:0")
n = lib.giga_debug_copy(net._engine().h, b"timeline", C.c_void_p(buf.data_ptr()), buf.numel(), C.c_void_p(0))
assert n > 0, lib.giga_last_error()
torch.cuda.synchronize()
t = buf[:n].cpu().numpy().view(np.uint64).reshape(-1, 32).astype(np.float64)
busy = t[t[:, 20] == t[:, 20].max()]   # CTAs with the most items (critical path)
m = np.median(busy, axis=0)
print(os.environ.get("GIGA_TIMELINE"), "CTAs", len(t), "items/CTA max", int(t[:, 20].max()), "(median over the busiest CTAs, cycles)")
print(f"  loader : wait-stage-free {m[0]:9.0f}  total {m[1]:9.0f}")
for k, nm in enumerate(("mma hi*hi", "mma lo*hi", "mma hi*lo")):
    print(f"  {nm}: wait-data {m[4+4*k]:9.0f}  wait-acc-free {m[5+4*k]:9.0f}  issue {m[6+4*k]:9.0f}  total {m[7+4*k]:9.0f}")
print(f"  drain  : wait-acc {m[16]:9.0f}  tmem-ld+add {m[17]:9.0f}  epilogue {m[18]:9.0f}  total {m[19]:9.0f}")
