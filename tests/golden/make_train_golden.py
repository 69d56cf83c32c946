#!/usr/bin/env python
"""Golden fixture for the training-step tail from the UNMODIFIED reference loss functions and torch.optim.Adam.

Run in the build container only:   python tests/golden/make_train_golden.py
scripts/train_giga.py imports ignite / tensorboard at module level (absent here), so the six loss functions (:161-195) are lifted out
of the reference script with `ast` and executed as they are; gradients with respect to the predictions come from autograd.
"""
import ast
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/scripts/train_giga.py"
WANT = {"loss_fn", "_qual_loss_fn", "_rot_loss_fn", "_quat_loss_fn", "_width_loss_fn", "_occ_loss_fn", "select"}


def reference_functions():
    import torch
    import torch.nn.functional as F
    src = open(REF).read()
    tree = ast.parse(src)
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in WANT]
    assert {n.name for n in body} == WANT
    ns = {"torch": torch, "F": F}
    exec(compile(ast.Module(body=body, type_ignores=[]), REF, "exec"), ns)
    return ns


def main():
    import torch
    from oracle import train_oracle as T
    ns = reference_functions()
    out = {}
    for tag, (B, M, seed) in {"a": (32, 2048, 0), "b": (5, 7, 1)}.items():
        arrs = T.seeded_batch(B, M, seed)
        t = [torch.from_numpy(a) for a in arrs]
        preds = [x.clone().requires_grad_(True) for x in t[:4]]
        loss, d = ns["loss_fn"](tuple(preds), tuple(t[4:]))
        loss.backward()
        out[f"{tag}_means"] = np.array([d[k].item() for k in ("loss_qual", "loss_rot", "loss_width", "loss_occ", "loss_all")], np.float64)
        for name, p in zip(("label", "rot", "width", "occ"), preds):
            out[f"{tag}_g_{name}"] = p.grad.numpy()
        out[f"{tag}_checksum"] = np.float64(sum(float(np.abs(a.astype(np.float64)).sum()) for a in arrs))
    # Adam: 12 steps of torch.optim.Adam on 1003 elements (not a multiple of 4), with and without weight decay
    rs = np.random.RandomState(9)
    p0 = rs.standard_normal(1003).astype(np.float32)
    gs = (rs.standard_normal((12, 1003)) * np.logspace(-4, 1, 1003)).astype(np.float32)
    out["adam_p0"], out["adam_g"] = p0, gs
    for tag, kw in {"plain": dict(lr=2e-4), "wd": dict(lr=1e-2, betas=(0.8, 0.95), eps=1e-6, weight_decay=0.1)}.items():
        p = torch.nn.Parameter(torch.from_numpy(p0.copy()))
        opt = torch.optim.Adam([p], **kw)
        traj = []
        for g in gs:
            p.grad = torch.from_numpy(g.copy())
            opt.step()
            traj.append(p.detach().numpy().copy())
        out[f"adam_{tag}"] = np.stack(traj)[[0, 5, 11]]
    np.savez_compressed(os.path.join(HERE, "train_golden.npz"), **out)
    print({k: getattr(v, "shape", v) for k, v in out.items()})
    print(out["a_means"], out["b_means"])


if __name__ == "__main__":
    main()
