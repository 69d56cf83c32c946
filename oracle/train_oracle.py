"""CPU oracle for the training-step tail (TEST INFRASTRUCTURE, not product).

numpy restatement of `loss_fn` (reference /root/reference/scripts/train_giga.py:161-195) with its analytic gradient, and of the Adam
update rule of the reference's optimizer (`torch.optim.Adam`, train_giga.py:67; third-party: torch, update rule of
torch/optim/adam.py `_single_tensor_adam`).  Pinned by tests/golden/train_golden.npz, made by tests/golden/make_train_golden.py from the
reference's own loss functions (extracted from the reference script by `ast`, executed unmodified, gradients by autograd) and from
torch.optim.Adam itself.
"""
from __future__ import annotations

import numpy as np


def _bce(x, t):
    """F.binary_cross_entropy(reduction='none'): logs clamped at -100 (aten/native/Loss.cpp)"""
    with np.errstate(divide="ignore"):
        return (t - 1) * np.maximum(np.log1p(-x.astype(np.float64)), -100) - t * np.maximum(np.log(x.astype(np.float64)), -100)


def _bce_grad(x, t):
    x = x.astype(np.float64)
    return (x - t) / np.maximum((1 - x) * x, 1e-12)


def loss(label_pred, rot_pred, width_pred, occ_pred, label, rotations, width, occ):
    """-> (loss_dict means [qual, rot, width, occ, all], gradients of loss_all w.r.t. the four predictions), float64 arithmetic"""
    f = lambda a: np.asarray(a, np.float64)
    lp, rp, wp, op, la, ro, wi, oc = (np.asarray(a, np.float32) for a in (label_pred, rot_pred, width_pred, occ_pred, label, rotations, width, occ))
    B, M = op.shape
    l_qual = _bce(lp, f(la))                                             # :177-178
    d = np.einsum("bk,bik->bi", f(rp), f(ro))                            # :187-188
    li = 1.0 - np.abs(d)
    l_rot = li.min(1)                                                    # :181-184
    dw = 40 * f(wp) - 40 * f(wi)
    l_width = dw * dw                                                    # :191-192
    l_occ = _bce(op, f(oc)).mean(-1)                                     # :194-195
    total = l_qual + f(la) * (l_rot + 0.01 * l_width) + l_occ            # :168
    means = np.array([l_qual.mean(), l_rot.mean(), l_width.mean(), l_occ.mean(), total.mean()])
    w0 = np.where(li[:, 0] < li[:, 1], 1.0, np.where(li[:, 0] == li[:, 1], 0.5, 0.0))
    g_rot = -(f(la) / B)[:, None] * (w0[:, None] * np.sign(d[:, :1]) * f(ro[:, 0]) + (1 - w0)[:, None] * np.sign(d[:, 1:]) * f(ro[:, 1]))
    grads = (_bce_grad(lp, f(la)) / B, g_rot, f(la) * 0.01 * 2 * dw * 40 / B, _bce_grad(op, f(oc)) / (B * M))
    return means, grads


def adam(param, grads, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
    """apply len(grads) Adam steps to `param` (float64 arithmetic) -> trajectory [steps][n]"""
    p = np.asarray(param, np.float64).copy()
    m = np.zeros_like(p)
    v = np.zeros_like(p)
    out = []
    for t, g in enumerate(grads, 1):
        g = np.asarray(g, np.float64)
        if weight_decay:
            g = g + weight_decay * p
        m = m + (1 - betas[0]) * (g - m)
        v = betas[1] * v + (1 - betas[1]) * g * g
        bc1, bc2 = 1 - betas[0] ** t, 1 - betas[1] ** t
        p = p - (lr / bc1) * m / (np.sqrt(v) / np.sqrt(bc2) + eps)
        out.append(p.copy())
    return np.stack(out)


def seeded_batch(B: int, M: int, seed: int = 0, edge_cases: bool = True):
    """predictions as `select(net(...))` produces them and targets as the data loader does (train_giga.py:140-158)"""
    rs = np.random.RandomState(3000 + seed)
    sig = lambda a: 1 / (1 + np.exp(-a))
    unit = lambda a: a / np.linalg.norm(a, axis=-1, keepdims=True)
    lp = sig(rs.standard_normal(B) * 2).astype(np.float32)
    rp = unit(rs.standard_normal((B, 4))).astype(np.float32)
    wp = (rs.random_sample(B) * 0.1).astype(np.float32)
    op = sig(rs.standard_normal((B, M)) * 3).astype(np.float32)
    la = (rs.random_sample(B) < 0.5).astype(np.float32)
    ro = unit(rs.standard_normal((B, 2, 4))).astype(np.float32)
    wi = (rs.random_sample(B) * 0.1).astype(np.float32)
    oc = (rs.random_sample((B, M)) < 0.3).astype(np.float32)
    if edge_cases and B >= 4 and M >= 4:
        lp[0], la[0] = 1.0, 0.0            # log(1 - x) = -inf -> clamp -100; backward denominator clamp 1e-12
        lp[1], la[1] = 0.0, 1.0
        op[2, 0], oc[2, 0] = 1.0, 0.0
        op[2, 1], oc[2, 1] = 0.0, 1.0
        op[2, 2], oc[2, 2] = 1.0, 1.0
        ro[3, 1] = -ro[3, 0]               # tie of the two rotation losses
        la[3] = 1.0
    return lp, rp, wp, op, la, ro, wi, oc
