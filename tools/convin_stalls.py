#!/usr/bin/env python
"""Debug: stall accounting of the tensor-core conv_in.  GIGA_TIMELINE=conv_in_tc python tools/convin_stalls.py"""
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
m = np.median(t, axis=0)
print("conv_in_tc CTAs", len(t), "(median cycles)")
print(f"  loader : wait-stage-free {m[0]:8.0f}  write+fence+arrive {m[1]:8.0f}  total {m[2]:8.0f}")
for k, nm in enumerate(("mma D1", "mma D2")):
    print(f"  {nm}: wait-data {m[4+4*k]:8.0f}  wait-acc-free {m[5+4*k]:8.0f}  issue {m[6+4*k]:8.0f}  total {m[7+4*k]:8.0f}")
print(f"  drain  : wait-acc {m[16]:8.0f}  tmem-ld {m[17]:8.0f}  reductions {m[18]:8.0f}  total {m[19]:8.0f}")
print(f"  reductions split (thread 0): bar1 {m[20]:8.0f}  relu+sts {m[21]:8.0f}  bar2 {m[22]:8.0f}  xz {m[23]:8.0f}  yz {m[24]:8.0f}")
st, en, sm = t[:, 28], t[:, 29], t[:, 30]
t0 = st.min()
print("kernel span us", (en.max() - t0) / 1e3, " CTA duration us median", np.median(en - st) / 1e3)
starts = np.sort((st - t0) / 1e3)
print("CTA start times us (every 32nd):", np.round(starts[::32], 1))
# co-residency: for each SM, max number of CTAs overlapping in time
mx = 0
for s_ in np.unique(sm):
    idx = np.where(sm == s_)[0]
    ev = sorted([(st[i], 1) for i in idx] + [(en[i], -1) for i in idx])
    c = 0
    for _, d in ev:
        c += d; mx = max(mx, c)
print("max CTAs resident on one SM at once:", mx)
