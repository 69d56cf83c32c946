"""TEST INFRASTRUCTURE -- the GIGA network in differentiable PyTorch ops on the GPU, and a bridge that pairs the library's forward with
PyTorch's own backward (ATen / cuDNN kernels).  Used as the SECOND gradient reference of tests/test_gpu_train_native.py (the gradient of a
ReLU / max-pool network is discontinuous: where the CPU oracle and a GPU forward decide a near-tie differently, the native backward must
agree with PyTorch's GPU autograd of the same function) and by tools/train_step_bench.py as the round-1 baseline.  Not part of the
product: giga_b200 contains no PyTorch compute."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

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
    """library forward + PyTorch backward; x / p / p_tsdf float32 tensors on the model's device"""
    named = [(k, v) for k, v in net.named_parameters()]
    if p is None:      # giga_geo: forward(inputs, p, p_tsdf) evaluates the TSDF head only
        return _Bridge.apply(net, x, p_tsdf, p_tsdf, [k for k, _ in named], *[v for _, v in named])
    return _Bridge.apply(net, x, p, p_tsdf, [k for k, _ in named], *[v for _, v in named])


