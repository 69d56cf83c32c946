#!/usr/bin/env python
"""Cost of re-committing the parameters after they changed (what an evaluation between training steps pays): device-side packer
(giga_ctx_commit_device, default for device-resident parameters) vs the host packer (giga_ctx_commit_params)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, giga_b200
from oracle import giga_oracle as O
for mode in ("auto", "host"):
    net = giga_b200.get_network("giga"); net.load_state_dict(O.seeded_state_dict(seed=1)); net = net.to("cuda:0")
    net._engine_raw().commit_mode = mode
    net._engine()
    ts = []
    for _ in range(10):
        with torch.no_grad(): net.decoder_width.fc_out.bias.add_(1e-3)
        torch.cuda.synchronize(); t0 = time.perf_counter(); net._engine(); torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
    print("re-commit (%s packer) ms: median %.3f min %.3f" % ("device" if mode == "auto" else "host", sorted(ts)[5], min(ts)))
