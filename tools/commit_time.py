#!/usr/bin/env python
"""Debug: cost of re-committing the parameters (what a training loop pays after every optimizer.step())."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, giga_b200
from oracle import giga_oracle as O
net = giga_b200.get_network("giga"); net.load_state_dict(O.seeded_state_dict(seed=1)); net = net.to("cuda:0")
eng = net._engine()
ts = []
for _ in range(10):
    with torch.no_grad(): net.decoder_width.fc_out.bias.add_(1e-3)
    torch.cuda.synchronize(); t0 = time.perf_counter(); net._engine(); torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
print("re-commit ms: median %.2f min %.2f" % (sorted(ts)[5], min(ts)))
