#!/usr/bin/env python
"""Debug: CTA start / end envelope of every persistent U-Net layer within one encode (globaltimer ns, relative to the first
layer's first CTA).  GIGA_LAYER_TIMES=1 [GIGA_TILE_DEPS=0|1] python tools/layer_times.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch, giga_b200
from giga_b200._lib import lib
from oracle import giga_oracle as O
B = int(os.environ.get("B", "32"))
net = giga_b200.get_network("giga"); net.load_state_dict(O.seeded_state_dict(seed=1)); net = net.to("cuda:0")
x = torch.rand(B, 40, 40, 40, device="cuda:0")
names = ["d0c1", "d0c2", "d1c1", "d1c2", "d2c1", "d2c2", "u0up", "u0c1", "u0c2", "u1up", "u1c1", "u1c2+fin"]
acc = []
for rep in range(8):
    net.encode_inputs(x)
    torch.cuda.synchronize()
    buf = torch.zeros(128, dtype=torch.float32, device="cuda:0")
    n = lib.giga_debug_copy(net._engine().h, b"layer_times", C.c_void_p(buf.data_ptr()), buf.numel(), C.c_void_p(0))
    assert n > 0, lib.giga_last_error()
    torch.cuda.synchronize()
    t = buf.cpu().numpy().view(np.uint64).reshape(16, 4)[:12].astype(np.float64)
    if rep >= 3: acc.append((t - t[0, 0]) / 1e3)
t = np.median(np.stack(acc), 0)
print(f"tile_deps={os.environ.get('GIGA_TILE_DEPS', '1')}  B={B}  (us since the first CTA of d0c1; median of 5 encodes)")
print(f"{'layer':<10}{'first start':>12}{'last start':>12}{'first end':>12}{'last end':>12}{'span':>9}")
for nme, r in zip(names, t):
    print(f"{nme:<10}{r[0]:12.1f}{r[1]:12.1f}{r[2]:12.1f}{r[3]:12.1f}{r[3] - r[0]:9.1f}")
print(f"U-Net total {t[11, 3] - t[0, 0]:.1f} us")
