"""CPU oracle for the VGN baseline network (TEST INFRASTRUCTURE, not product).

A functional fp32 restatement of the reference `ConvNet` (reference paths relative to /root/reference/src/vgn): networks.py:48-63
(forward), :172-188 (Encoder), :191-212 (Decoder), :37-45 (conv / conv_stride: padding k // 2, stride 1 / 2).  The reference is PyTorch, so
the restatement calls the ATen conv3d it calls; F.interpolate(x, size) (default mode 'nearest', size an exact multiple) is restated as
the index map src = dst // 2.  Pinned against the reference's own `ConvNet` (imported in the build container by
tests/golden/make_vgn_golden.py; fixture tests/golden/vgn_golden.npz; checked by tests/test_vgn_oracle.py).
Everything takes a flat state_dict mapping with the reference's parameter names.
"""
from __future__ import annotations

from typing import Dict, Mapping

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
LAYERS = (("encoder.conv1", 1, 16, 5), ("encoder.conv2", 16, 32, 3), ("encoder.conv3", 32, 64, 3),
          ("decoder.conv1", 64, 64, 3), ("decoder.conv2", 64, 32, 3), ("decoder.conv3", 32, 16, 5),
          ("conv_qual", 16, 1, 5), ("conv_rot", 16, 4, 5), ("conv_width", 16, 1, 5))


def _conv(sd, name, x, stride=1):
    w = sd[name + ".weight"]
    return F.conv3d(x, w, sd[name + ".bias"], stride=stride, padding=w.shape[-1] // 2)      # networks.py:37-45


def _up2(x: Tensor) -> Tensor:
    """F.interpolate(x, 2 * D) nearest: out[..., i, j, k] = x[..., i // 2, j // 2, k // 2] (networks.py:203,207,211)"""
    idx = torch.arange(2 * x.shape[-1]) // 2
    return x[:, :, idx][:, :, :, idx][:, :, :, :, idx]


def encoder(sd: Mapping[str, Tensor], x: Tensor) -> Tensor:
    """networks.py:179-188"""
    for i in (1, 2, 3):
        x = F.relu(_conv(sd, f"encoder.conv{i}", x, stride=2))
    return x


def decoder(sd: Mapping[str, Tensor], x: Tensor) -> Tensor:
    """networks.py:198-212"""
    for i in (1, 2, 3):
        x = _up2(F.relu(_conv(sd, f"decoder.conv{i}", x)))
    return x


def forward(sd: Mapping[str, Tensor], x: Tensor, capture: dict = None):
    """networks.py:57-63: x (B,1,40,40,40) -> qual (B,1,40,40,40), rot (B,4,40,40,40), width (B,1,40,40,40)"""
    e = encoder(sd, x)
    d = decoder(sd, e)
    if capture is not None:
        capture["enc"], capture["dec"] = e, d
    qual = torch.sigmoid(_conv(sd, "conv_qual", d))
    rot = F.normalize(_conv(sd, "conv_rot", d), dim=1)
    width = _conv(sd, "conv_width", d)
    return qual, rot, width


def seeded_state_dict(seed: int = 1, gain: float = 1.0, width_bias: float = 5.0) -> Dict[str, Tensor]:
    """He-scaled normal weights (activations stay O(1) through the six ReLU layers), N(0, 0.1^2) biases, from numpy's frozen legacy
    MT19937 stream in state_dict order; conv_width.bias is shifted so predicted widths straddle the planner's 1.33..9.33-voxel gate."""
    rs = np.random.RandomState(seed)
    sd = {}
    for name, ci, co, k in LAYERS:
        std = gain * np.sqrt(2.0 / (ci * k ** 3))
        sd[name + ".weight"] = torch.from_numpy((rs.standard_normal((co, ci, k, k, k)) * std).astype(np.float32))
        sd[name + ".bias"] = torch.from_numpy((rs.standard_normal((co,)) * 0.1).astype(np.float32))
    sd["conv_width.bias"] = sd["conv_width.bias"] + width_bias
    sd["conv_qual.weight"] = sd["conv_qual.weight"] * 4.0     # spread the quality logits so some voxels pass the 0.9 threshold
    sd["conv_width.weight"] = sd["conv_width.weight"] * 6.0   # ... and the widths beyond both ends of the gate
    return sd


def seeded_inputs(B: int, seed: int = 0) -> Tensor:
    """TSDF-like volumes (B,1,40,40,40) in [0,1]: smooth blobs plus noise, empty (0) regions like unobserved space."""
    rs = np.random.RandomState(2000 + seed)
    g = np.stack(np.meshgrid(*[np.arange(40, dtype=np.float32)] * 3, indexing="ij"), -1)
    out = np.zeros((B, 1, 40, 40, 40), np.float32)
    for b in range(B):
        v = np.ones((40, 40, 40), np.float32)
        for _ in range(5):
            c = rs.uniform(6, 34, 3).astype(np.float32)
            r = rs.uniform(4, 9)
            v = np.minimum(v, np.abs(np.linalg.norm(g - c, axis=-1) - r) / 4.0)
        v = np.clip(v + rs.standard_normal(v.shape).astype(np.float32) * 0.02, 0, 1)
        v[v > 0.98] = 0.0
        out[b, 0] = v
    return torch.from_numpy(out)
