#!/usr/bin/env python
"""Debug: stall accounting of the tensor-core conv_in kernel.  GIGA_TIMELINE=conv_in_tc python tools/convin_stalls.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch, giga_b200
from giga_b200._lib import lib
from oracle import giga_oracle as O
B = int(os.environ.get("B", "32"))
net = giga_b200.get_network("giga"); net.load_state_dict(O.seeded_state_dict(seed=1)); net = net.to("cuda:0")
x = torch.rand(B, 40, 40, 40, device="cuda:0")
for _ in range(3): net.encode_inputs(x)
torch.cuda.synchronize()
buf = torch.zeros(1 << 20, dtype=torch.float32, device="cuda:0")
n = lib.giga_debug_copy(net._engine().h, b"timeline", C.c_void_p(buf.data_ptr()), buf.numel(), C.c_void_p(0))
assert n > 0, lib.giga_last_error()
torch.cuda.synchronize()
t = buf[:n].cpu().numpy().view(np.uint64).reshape(-1, 32).astype(np.float64)
m = np.median(t, axis=0)
print(f"conv_in_tc CTAs {len(t)} (median cycles over CTAs, 40 march steps)")
print(f"  load+mma warp: wait-slot-free {m[0]:8.0f}  wait-plane {m[1]:8.0f}  wait-acc-free {m[2]:8.0f}  issue {m[3]:8.0f}  total {m[4]:8.0f}")
print(f"  drain warps  : wait-acc {m[8]:8.0f}  barrier-A {m[9]:8.0f}  tmem-ld+relu+sts {m[10]:8.0f}  barrier-B {m[11]:8.0f}  reductions {m[12]:8.0f}  total {m[13]:8.0f}")
