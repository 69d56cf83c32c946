"""Training on the library's own kernels: the native training step, the fused loss / flat Adam tail, the gradient exchange.

  * `native_forward` / `_NativeStep` (csrc/train_bwd.cuh, giga_train_forward / giga_train_backward): what `net(...)` runs when gradient
    mode is on and parameters require gradients -- replaces the autograd graph of scripts/train_giga.py:199-211 (`_update`).  Forward on
    the fp32 kernels from the live parameter tensors (packed on the device), backward = hand-written kernels for every layer; gradients
    are accumulated either straight into `.grad` (parameters owned by `Adam` below: one flat buffer) or handed to autograd.
  * `select(out)` (train_giga.py:153-158), `loss_fn(y_pred, y)` (:161-174) with the reference's signatures and return values -- ONE CUDA
    launch for the loss terms and the gradient of loss.mean() w.r.t. the predictions (giga_loss) -- and `Adam`, torch.optim.Adam's
    constructor for the arguments train_giga.py uses (:67): parameters, gradients and moments in four flat buffers, ONE launch per step
    (giga_adam_step); `Adam.allreduce_gradients()` = one in-place NCCL all-reduce of the flat 581,863-element gradient buffer
    (SURVEY.md section 8e), `allreduce_gradients(params)` the same for a plain torch optimizer.
  * gradients with respect to the grasp query positions (`grad_refine`, models/__init__.py:136-164) come from the same backward kernels
    (fc_p plus grid_sampler's grid gradient).  There is no PyTorch / ATen compute anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check, lib

# ------------------------------------------------------------------------------------------------------
# the native training step (csrc/train_bwd.cuh): differentiable forward + hand-written backward kernels
# ------------------------------------------------------------------------------------------------------
class _NativeStep(torch.autograd.Function):
    """forward: giga_train_forward (fp32 FMA-pipe kernels on the live parameter tensors, activations kept in the engine);
    backward: giga_train_backward (decoder / grid-sample / conv / transpose-conv / pool / conv_in backward kernels).  Parameter
    gradients are accumulated by the kernels either straight into `.grad` (parameters owned by `training.Adam`: one flat buffer) or
    into a fresh flat buffer whose views are handed to autograd."""

    @staticmethod
    def forward(ctx, net, x, p, p_tsdf, names, *params):
        from .model import _prep, GRID
        eng = net._engine_raw()
        dev = eng.device
        x = _prep(x, dev)
        if x.dim() != 4 or tuple(x.shape[1:]) != (GRID, GRID, GRID):
            raise _lib.GigaError(f"inputs must be (B,{GRID},{GRID},{GRID}), got {tuple(x.shape)}")
        B = x.shape[0]

        def pts(t, what):
            if t is None:
                return None, 0
            t = _prep(t, dev)
            if t.dim() != 3 or t.shape[2] != 3 or t.shape[0] != B:
                raise _lib.GigaError(f"{what} must be (B,N,3) with B={B}, got {tuple(t.shape)}")
            return t, t.shape[1]

        ctx.p_grad = bool(p is not None and p.requires_grad)
        if p_tsdf is not None and p_tsdf.requires_grad:
            raise _lib.GigaError("gradients with respect to p_tsdf are not provided (no caller of the reference asks for them)")
        p, Ng = pts(p, "points")
        p_tsdf, No = pts(p_tsdf, "p_tsdf")
        if p_tsdf is not None and not hasattr(net, "decoder_tsdf"):
            raise AttributeError("this model has no decoder_tsdf")      # as the reference (models/__init__.py:64)
        if any(q.device != dev or q.dtype != torch.float32 or not q.is_contiguous() for q in params):
            raise _lib.GigaError("the native training step takes contiguous fp32 parameters on the model's device")
        direct = all(getattr(q, "_giga_direct", False) and q.grad is not None and q.grad.is_contiguous() and q.grad.data_ptr() % 16 == 0
                     for q in params if q.requires_grad)
        flat = None
        if direct:
            grads = [q.grad if q.requires_grad else None for q in params]
        else:
            sizes = [(q.numel() + 3) // 4 * 4 for q in params]
            flat = torch.zeros(sum(sizes), device=dev, dtype=torch.float32)
            grads, off = [], 0
            for q, m in zip(params, sizes):
                grads.append(flat[off:off + q.numel()].view(q.shape))
                off += m
        frozen = None
        if any(g is None for g in grads):     # parameters with requires_grad = False still need somewhere to write
            frozen = torch.zeros(max(q.numel() for q in params), device=dev, dtype=torch.float32)
            grads = [g if g is not None else frozen for g in grads]
        n = len(params)
        key = (tuple(q.data_ptr() for q in params), tuple(g.data_ptr() for g in grads))
        if eng.__dict__.get("_train_key") != key:
            c_names = (C.c_char_p * n)(*[k.encode() for k in names])
            vals = (C.c_void_p * n)(*[q.data_ptr() for q in params])
            gptr = (C.c_void_p * n)(*[g.data_ptr() for g in grads])
            check(lib.giga_train_bind(eng.h, n, c_names, vals, gptr), "giga_train_bind")
            eng.__dict__["_train_key"] = key
        mk = lambda *s: torch.empty(s, device=dev, dtype=torch.float32)
        qual = rot = width = occ = None
        if p is not None:
            qual, rot, width = mk(B, Ng), mk(B, Ng, 4), mk(B, Ng)
        if p_tsdf is not None:
            occ = mk(B, No)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        check(lib.giga_train_forward(eng.h, _ptr(x), B, _ptr(p), Ng, _ptr(p_tsdf), No, int(bool(getattr(net, "detach_tsdf", False))),
                                     _ptr(qual), _ptr(rot), _ptr(width), _ptr(occ), st), "giga_train_forward")
        eng.__dict__["_train_token"] = eng.__dict__.get("_train_token", 0) + 1
        ctx.eng, ctx.token, ctx.direct, ctx.grads, ctx.flat, ctx.frozen = eng, eng.__dict__["_train_token"], direct, grads, flat, frozen
        ctx.keep = (x, p, p_tsdf)       # the backward kernels re-read the inputs
        ctx.slots = [o is not None for o in (qual, rot, width, occ)]
        ctx.params_rg = [q.requires_grad for q in params]
        return tuple(o for o in (qual, rot, width, occ) if o is not None)

    @staticmethod
    def backward(ctx, *gouts):
        eng = ctx.eng
        if eng.__dict__.get("_train_token") != ctx.token:
            raise _lib.GigaError("backward() of a forward that is no longer the engine's latest training forward "
                                 "(one giga_b200 model keeps the activations of ONE forward at a time)")
        dev = eng.device
        it = iter(gouts)
        gs = []
        for present in ctx.slots:
            g = next(it) if present else None
            gs.append(g.to(device=dev, dtype=torch.float32).contiguous() if g is not None else None)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        gp = torch.zeros_like(ctx.keep[1]) if ctx.p_grad else None
        check(lib.giga_train_backward(eng.h, _ptr(gs[0]), _ptr(gs[1]), _ptr(gs[2]), _ptr(gs[3]), _ptr(gp), st), "giga_train_backward")
        eng.__dict__["_train_token"] += 1      # consumed
        head = (None, None, gp, None, None)
        if ctx.direct:
            return head + (None,) * len(ctx.grads)
        return head + tuple(g if rg else None for g, rg in zip(ctx.grads, ctx.params_rg))


def native_forward(net, x, p, p_tsdf):
    """differentiable forward through the native training step; returns the tuple forward() returns"""
    named = list(net.named_parameters())
    return _NativeStep.apply(net, x, p, p_tsdf, [k for k, _ in named], *[v for _, v in named])


# ------------------------------------------------------------------------------------------------------
# data-parallel gradient exchange (one process per GPU, scenes sharded by the data loader)
# ------------------------------------------------------------------------------------------------------
def allreduce_gradients(params: Sequence[torch.nn.Parameter], group=None, average: bool = True) -> Optional[torch.Tensor]:
    """One flat all-reduce (sum, then /world) over every parameter gradient: 581,863 floats = 2.33 MB for GIGA,
    latency-bound on NVLink -> a single bucket.  Parameters without a gradient
    contribute zeros (all ranks must agree on the layout).  Returns the flat buffer (or None single-process)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    params = list(params)
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off:off + n].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
    return flat


# ------------------------------------------------------------------------------------------------------
# the native training-step tail: fused loss (value + gradient) and flat Adam (csrc/train.cuh)
# ------------------------------------------------------------------------------------------------------
_ENGINES = {}


def _engine(device: torch.device):
    from .model import _Engine
    if device.type != "cuda":
        raise _lib.GigaError("giga_b200 runs on CUDA (sm_100a) only; there is no CPU path")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _ENGINES:
        _ENGINES[idx] = _Engine(torch.device("cuda", idx))
    return _ENGINES[idx]


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def select(out):
    """train_giga.py:153-158"""
    qual_out, rot_out, width_out, occ = out
    return qual_out.squeeze(-1), rot_out.squeeze(1), width_out.squeeze(-1), torch.sigmoid(occ)


class _FusedLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, label_pred, rot_pred, width_pred, occ_pred, label, rotations, width, occ):
        dev = label_pred.device
        eng = _engine(dev)
        f = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        lp, rp, wp, op, la, ro, wi, oc = (f(t) for t in (label_pred, rot_pred, width_pred, occ_pred, label, rotations, width, occ))
        B = lp.numel()
        if rp.shape != (B, 4) or wp.numel() != B or ro.shape != (B, 2, 4) or op.dim() != 2 or op.shape[0] != B or oc.shape != op.shape \
                or la.numel() != B or wi.numel() != B:
            raise _lib.GigaError("loss_fn: label_pred/width_pred (B,), rotation_pred (B,4), occ_pred (B,M); label/width (B,), rotations (B,2,4), occ (B,M)")
        M = op.shape[1]
        out = torch.empty(5, device=dev, dtype=torch.float32)
        grads = [torch.empty_like(t) for t in (lp, rp, wp, op)]
        check(lib.giga_loss(eng.h, _ptr(lp), _ptr(rp), _ptr(wp), _ptr(op), _ptr(la), _ptr(ro), _ptr(wi), _ptr(oc), B, M, _ptr(out),
                            *[_ptr(g) for g in grads], C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "giga_loss")
        ctx.save_for_backward(*grads)
        ctx.shapes = [t.shape for t in (label_pred, rot_pred, width_pred, occ_pred)]
        return out

    @staticmethod
    def backward(ctx, g):
        # only loss_all (element 4) is a training objective; the other four entries are the detached-by-use loss_dict means
        scale = g[4]
        return tuple((gr * scale).reshape(sh) for gr, sh in zip(ctx.saved_tensors, ctx.shapes)) + (None,) * 4


def loss_fn(y_pred, y):
    """train_giga.py:161-174 -> (loss, loss_dict)"""
    label_pred, rotation_pred, width_pred, occ_pred = y_pred
    label, rotations, width, occ = y
    out = _FusedLoss.apply(label_pred, rotation_pred, width_pred, occ_pred, label, rotations, width, occ)
    d = out.detach()
    loss_dict = {"loss_qual": d[0], "loss_rot": d[1], "loss_width": d[2], "loss_occ": d[3], "loss_all": d[4]}
    return out[4], loss_dict


class Adam(torch.optim.Optimizer):
    """torch.optim.Adam(params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0) for CUDA fp32 parameters, one launch per step.
    The parameters are re-pointed at one flat buffer at construction (their values are kept; `.data` stays a live view, so the model sees
    every update), and their `.grad` at a flat gradient buffer: backward() accumulates straight into it and zero_grad() is one memset."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameter")     # torch/optim/adam.py:52-62
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._flat = []
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.requires_grad]
            if not ps:
                continue
            dev = ps[0].device
            if dev.type != "cuda" or any(p.device != dev or p.dtype != torch.float32 for p in ps):
                raise _lib.GigaError("giga_b200.training.Adam takes fp32 parameters on one CUDA device")
            n = sum((p.numel() + 3) // 4 * 4 for p in ps)
            buf = {k: torch.zeros(n, device=dev, dtype=torch.float32) for k in ("param", "grad", "exp_avg", "exp_avg_sq")}
            off = 0
            with torch.no_grad():
                for p in ps:
                    m = p.numel()
                    view = buf["param"][off:off + m].view(p.shape)
                    view.copy_(p)
                    p.data = view
                    p.grad = buf["grad"][off:off + m].view(p.shape)
                    p._giga_direct = True      # the native backward accumulates straight into this flat gradient buffer
                    off += (m + 3) // 4 * 4
            self._flat.append((group, buf, n, dev))
        self._steps = 0

    def allreduce_gradients(self, group=None, average: bool = True):
        """data-parallel exchange on the flat gradient buffer itself: one all-reduce per parameter group, in place"""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        for _, buf, _, _ in self._flat:
            dist.all_reduce(buf["grad"], op=dist.ReduceOp.SUM, group=group)
            if average:
                buf["grad"] /= dist.get_world_size(group)

    def zero_grad(self, set_to_none: bool = False):
        """keeps the flat gradient buffer bound (set_to_none would detach the views)"""
        for _, buf, _, _ in self._flat:
            buf["grad"].zero_()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self._steps += 1
        for group, buf, n, dev in self._flat:
            for p in group["params"]:
                if p.requires_grad and (p.grad is None or p.grad.data_ptr() < buf["grad"].data_ptr()
                                        or p.grad.data_ptr() >= buf["grad"].data_ptr() + 4 * n):
                    raise _lib.GigaError("a parameter's .grad no longer views the flat gradient buffer (use this optimizer's zero_grad())")
            b1, b2 = group["betas"]
            check(lib.giga_adam_step(_engine(dev).h, _ptr(buf["param"]), _ptr(buf["grad"]), _ptr(buf["exp_avg"]), _ptr(buf["exp_avg_sq"]), n,
                                     self._steps, float(group["lr"]), float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]),
                                     C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "giga_adam_step")
            # the kernel wrote through raw pointers: bump the autograd version counters so that readers which key on them (this package's
            # parameter cache, autograd's saved-tensor checks) see the update
            torch.autograd.graph.increment_version([p for p in group["params"] if p.requires_grad])
        return loss
