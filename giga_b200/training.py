"""Training on the library's own kernels: the native training step, the fused loss / flat Adam tail, the gradient exchange -- and the
round-1 recompute bridge, kept only for gradients with respect to query positions.

  * `native_forward` / `_NativeStep` (csrc/train_bwd.cuh, giga_train_forward / giga_train_backward): what `net(...)` runs when gradient
    mode is on and parameters require gradients -- replaces the autograd graph of scripts/train_giga.py:199-211 (`_update`).  Forward on
    the fp32 kernels from the live parameter tensors (packed on the device), backward = hand-written kernels for every layer; gradients
    are accumulated either straight into `.grad` (parameters owned by `Adam` below: one flat buffer) or handed to autograd.
  * `select(out)` (train_giga.py:153-158), `loss_fn(y_pred, y)` (:161-174) with the reference's signatures and return values -- ONE CUDA
    launch for the loss terms and the gradient of loss.mean() w.r.t. the predictions (giga_loss) -- and `Adam`, torch.optim.Adam's
    constructor for the arguments train_giga.py uses (:67): parameters, gradients and moments in four flat buffers, ONE launch per step
    (giga_adam_step); `Adam.allreduce_gradients()` = one in-place NCCL all-reduce of the flat 581,863-element gradient buffer
    (SURVEY.md section 8e), `allreduce_gradients(params)` the same for a plain torch optimizer.
  * `bridged_forward` / `_Bridge` (opt-in, `net.enable_training_bridge()`): forward by the CUDA library, backward = PyTorch autograd
    recompute on the GPU.  Used by `grad_refine` (gradients w.r.t. the query positions) and by the tests as a second
    gradient reference; never on the training path, never on the CPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import _lib
from ._lib import check, lib

PLANES = ("xz", "xy", "yz")
_AX = {"xz": (0, 2), "xy": (0, 1), "yz": (1, 2)}


# ------------------------------------------------------------------------------------------------------
# the model function in differentiable PyTorch ops (GPU), parameterised by a name -> tensor mapping
# ------------------------------------------------------------------------------------------------------
def _unet(sd: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    g = lambda n: sd["encoder.unet." + n]
    enc = []
    for i in range(3):
        x = F.relu(F.conv2d(x, g(f"down_convs.{i}.conv1.weight"), g(f"down_convs.{i}.conv1.bias"), padding=1))
        x = F.relu(F.conv2d(x, g(f"down_convs.{i}.conv2.weight"), g(f"down_convs.{i}.conv2.bias"), padding=1))
        enc.append(x)
        if i < 2:
            x = F.max_pool2d(x, 2, 2)
    for i in range(2):
        up = F.conv_transpose2d(x, g(f"up_convs.{i}.upconv.weight"), g(f"up_convs.{i}.upconv.bias"), stride=2)
        x = torch.cat((up, enc[-(i + 2)]), 1)
        x = F.relu(F.conv2d(x, g(f"up_convs.{i}.conv1.weight"), g(f"up_convs.{i}.conv1.bias"), padding=1))
        x = F.relu(F.conv2d(x, g(f"up_convs.{i}.conv2.weight"), g(f"up_convs.{i}.conv2.bias"), padding=1))
    return F.conv2d(x, g("conv_final.weight"), g("conv_final.bias"))


def _conv_in(x, w, b):
    """Conv3d(1, 32, 3, padding=1) as 27 shifted views x one matmul.  Same arithmetic as F.conv3d; cuDNN's weight-gradient kernel for a
    single input channel (wgrad2d_grouped_direct) took 37 ms per step at batch 64 -- 58 % of the whole training step -- while this
    form differentiates into a [32 x 27] <- [32 x B*64000] x [B*64000 x 27] matmul."""
    B = x.shape[0]
    xp = F.pad(x, (1, 1, 1, 1, 1, 1))
    cols = torch.stack([xp[:, dx:dx + 40, dy:dy + 40, dz:dz + 40] for dx in range(3) for dy in range(3) for dz in range(3)], 1)   # [B,27,40,40,40]
    f = torch.matmul(w.reshape(32, 27), cols.reshape(B, 27, -1)) + b.view(1, 32, 1)
    return f.view(B, 32, 40, 40, 40)


def _encode(sd, x):
    f = F.relu(_conv_in(x, sd["encoder.conv_in.weight"], sd["encoder.conv_in.bias"]))  # [b,c,ix,iy,iz]
    # 40^3 voxels onto 40^2 cells: the scatter_mean is the mean along the perpendicular axis
    pre = {"xz": f.mean(3).transpose(2, 3), "xy": f.mean(4).transpose(2, 3), "yz": f.mean(2).transpose(2, 3)}
    return {k: _unet(sd, v) for k, v in pre.items()}


def _norm_axis(v):
    t = v / 1.00001 + 0.5
    t = torch.where(t >= 1, torch.full_like(t, 1 - 10e-6), t)
    return torch.where(t < 0, torch.zeros_like(t), t)


def _features(p, planes):
    out = []
    for k in PLANES:
        a0, a1 = _AX[k]
        uv = torch.stack((_norm_axis(p[..., a0]), _norm_axis(p[..., a1])), -1)
        grid = (2.0 * uv - 1.0)[:, :, None]
        out.append(F.grid_sample(planes[k], grid, padding_mode="border", align_corners=True, mode="bilinear").squeeze(-1))
    return torch.cat(out, 1).transpose(1, 2)


def _head(sd, name, p, c):
    pre = f"decoder_{name}."
    net = F.linear(p, sd[pre + "fc_p.weight"], sd[pre + "fc_p.bias"])
    for i in range(5):
        net = net + F.linear(c, sd[pre + f"fc_c.{i}.weight"], sd[pre + f"fc_c.{i}.bias"])
        h = F.linear(F.relu(net), sd[pre + f"blocks.{i}.fc_0.weight"], sd[pre + f"blocks.{i}.fc_0.bias"])
        net = net + F.linear(F.relu(h), sd[pre + f"blocks.{i}.fc_1.weight"], sd[pre + f"blocks.{i}.fc_1.bias"])
    return F.linear(F.relu(net), sd[pre + "fc_out.weight"], sd[pre + "fc_out.bias"]).squeeze(-1)


def _forward_torch(sd, x, p, p_tsdf, detach_tsdf: bool, has_grasp: bool):
    planes = _encode(sd, x)
    outs = []
    if has_grasp:
        c = _features(p, planes)
        outs += [torch.sigmoid(_head(sd, "qual", p, c)), F.normalize(_head(sd, "rot", p, c), dim=2), _head(sd, "width", p, c)]
    if p_tsdf is not None:
        pl = {k: v.detach() for k, v in planes.items()} if detach_tsdf else planes
        outs.append(_head(sd, "tsdf", p_tsdf, _features(p_tsdf, pl)))
    return tuple(outs)


class _Bridge(torch.autograd.Function):
    """forward: native CUDA kernels; backward: autograd through `_forward_torch` (recompute)."""

    @staticmethod
    def forward(ctx, net, x, p, p_tsdf, names, *params):
        with torch.no_grad():
            outs = net._forward_native(x, p, p_tsdf)
        ctx.net, ctx.names, ctx.has_tsdf = net, names, p_tsdf is not None
        ctx.p_grad = bool(p.requires_grad)
        ctx.save_for_backward(x, p, p_tsdf if p_tsdf is not None else x.new_empty(0), *params)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        x, p, pt, *params = ctx.saved_tensors
        leaves = [t.detach().requires_grad_(True) for t in params]
        sd = dict(zip(ctx.names, leaves))
        # the recompute runs under the caller's torch.backends flags, exactly like the reference's own training step would
        # (PyTorch's default lets cuDNN use TF32 for convolutions; set torch.backends.cudnn.allow_tf32 = False for fp32 gradients)
        return _Bridge._backward_impl(ctx, leaves, sd, x, p, pt, grads)

    @staticmethod
    def _backward_impl(ctx, leaves, sd, x, p, pt, grads):
        with torch.enable_grad():
            p_leaf = p.detach().requires_grad_(True) if ctx.p_grad else p      # grad_refine differentiates w.r.t. the query positions
            outs = _forward_torch(sd, x, p_leaf, pt if ctx.has_tsdf else None, getattr(ctx.net, "detach_tsdf", False), hasattr(ctx.net, "decoder_qual"))
            pairs = [(o, g) for o, g in zip(outs, grads) if g is not None]
            wrt = list(leaves) + ([p_leaf] if ctx.p_grad else [])
            gp = torch.autograd.grad([o for o, _ in pairs], wrt, [g for _, g in pairs], allow_unused=True)
        gpos = gp[-1] if ctx.p_grad else None
        return (None, None, gpos, None, None) + tuple(gp[:len(leaves)])


def bridged_forward(net, x, p, p_tsdf):
    named = [(k, v) for k, v in net.named_parameters()]
    return _Bridge.apply(net, x, p, p_tsdf, [k for k, _ in named], *[v for _, v in named])


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
        check(lib.giga_train_backward(eng.h, _ptr(gs[0]), _ptr(gs[1]), _ptr(gs[2]), _ptr(gs[3]), st), "giga_train_backward")
        eng.__dict__["_train_token"] += 1      # consumed
        if ctx.direct:
            return (None,) * 5 + (None,) * len(ctx.grads)
        return (None,) * 5 + tuple(g if rg else None for g, rg in zip(ctx.grads, ctx.params_rg))


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
