#!/usr/bin/env python
"""Per-parameter gradient error of the native training step against CPU autograd through the oracle (diagnostic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from oracle import giga_oracle as O
from tests.util import make_net
from tests.test_gpu_train_native import _batch, _loss

B, No = int(os.environ.get("B", 4)), int(os.environ.get("NO", 128))
sd = O.seeded_state_dict(seed=1)
net = make_net("giga", sd, frozen=False)
x, p, pt, y = _batch(B, No, seed=50 + B)
out = net(x.cuda(), p.cuda(), p_tsdf=pt.cuda())
ref_leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
ref_out = O.forward(ref_leaves, x, p, pt)
for a, b, nm in zip(out, ref_out, ("qual", "rot", "width", "occ")):
    print(f"forward {nm}: max abs err {(a.detach().cpu() - b.detach()).abs().max().item():.3e}")
_loss(out, y, "cuda").backward()
torch.cuda.synchronize()
_loss(ref_out, y, "cpu").backward()
for k, prm in net.named_parameters():
    r = ref_leaves[k].grad
    g = prm.grad.detach().cpu()
    err = ((g - r).abs().max() / (r.abs().max() + 1e-12)).item()
    flag = "" if err < 2e-4 else "   <-- BAD"
    print(f"{k:55s} ref max {r.abs().max().item():.3e}  got max {g.abs().max().item():.3e}  rel err {err:.2e}{flag}")

if os.environ.get("BRIDGE"):
    # the same gradients from PyTorch's own GPU kernels (the opt-in bridge, TF32 off): how far do two correct fp32 implementations
    # differ on this batch (ReLU / max-pool decisions at near-ties are discontinuities of the gradient)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from tests.torch_bridge import bridged_forward
    net2 = make_net("giga", sd, frozen=False)
    out2 = bridged_forward(net2, x.cuda(), p.cuda(), pt.cuda())
    _loss(out2, y, "cuda").backward()
    torch.cuda.synchronize()
    print("---- bridge (ATen/cuDNN on the GPU) vs CPU oracle, and native vs bridge ----")
    for (k, prm), (_, prm2) in zip(net.named_parameters(), net2.named_parameters()):
        r = ref_leaves[k].grad
        g, g2 = prm.grad.detach().cpu(), prm2.grad.detach().cpu()
        e_b = ((g2 - r).abs().max() / (r.abs().max() + 1e-12)).item()
        e_nb = ((g - g2).abs().max() / (r.abs().max() + 1e-12)).item()
        if max(e_b, e_nb) > 2e-4:
            print(f"{k:55s} bridge-vs-cpu {e_b:.2e}   native-vs-bridge {e_nb:.2e}")
