"""CPU oracle for GIGA's dense-inference hot path (TEST INFRASTRUCTURE, not product).

A plain PyTorch-fp32, CPU-only, functional restatement of the reference path
`ConvolutionalOccupancyNetwork.forward` (reference file:line cited per function,
paths relative to /root/reference/src/vgn/ConvONets).  The reference itself is
Python/PyTorch, so the restatement calls the same ATen operators the reference
calls (conv3d, conv2d, conv_transpose2d, max_pool2d, grid_sample, linear); the
third-party `torch_scatter.scatter_mean` (torch-scatter 2.0.6, pinned in
environment.yaml:145, NOT vendored under /root/reference) is restated from its
published semantics (out += scatter_sum(src); out /= clamp(count, 1)).

Pinning: the reference ships no tests / golden vectors for this path, so this
oracle is pinned against outputs of the reference ITSELF, imported in the build
container by `tests/golden/make_golden.py` (script + small fixtures committed
under tests/golden/).  `tests/test_oracle_golden.py` checks every stage.

Everything takes a flat ``state_dict``-style mapping ``name -> tensor`` using the
reference's own parameter names.
"""
from __future__ import annotations

from typing import Dict, Mapping, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
PLANES = ("xz", "xy", "yz")
HEADS = ("qual", "rot", "width", "tsdf")
_PLANE_AXES = {"xz": (0, 2), "xy": (0, 1), "yz": (1, 2)}


# --------------------------------------------------------------------------
# common.py
# --------------------------------------------------------------------------
def normalize_coordinate(p: Tensor, padding: float = 0.0, plane: str = "xz") -> Tensor:
    """common.py:238-261 -- pick the plane's two axes, /(1+padding+10e-6), +0.5,
    then the two one-sided clamps (>=1 -> 1-10e-6 ; <0 -> 0)."""
    a0, a1 = _PLANE_AXES[plane]
    xy = p[:, :, [a0, a1]]
    xy_new = xy / (1 + padding + 10e-6)
    xy_new = xy_new + 0.5
    # reference does the masked assignment only when max()/min() trips; the
    # element-wise result is identical to an unconditional masked assignment.
    xy_new = torch.where(xy_new >= 1, torch.full_like(xy_new, 1 - 10e-6), xy_new)
    xy_new = torch.where(xy_new < 0, torch.zeros_like(xy_new), xy_new)
    return xy_new


def coordinate2index(x: Tensor, reso: int) -> Tensor:
    """common.py:303-318 (coord_type='2d') -- floor(x*reso); idx = x0 + reso*x1."""
    xi = (x * reso).long()
    index = xi[:, :, 0] + reso * xi[:, :, 1]
    return index[:, None, :]


def scatter_mean(src: Tensor, index: Tensor, out: Tensor) -> Tensor:
    """torch_scatter 2.0.6 `scatter_mean(src, index, dim=-1, out=out)` restated:
    sum into `out`, count per slot clamped to >=1, true-divide.  Call site:
    encoder/voxels.py:63-66."""
    idx = index.expand_as(src)
    out.scatter_add_(-1, idx, src)
    count = torch.zeros_like(out)
    count.scatter_add_(-1, idx, torch.ones_like(src))
    count.clamp_(min=1)
    out.div_(count)
    return out


# --------------------------------------------------------------------------
# encoder/unet.py
# --------------------------------------------------------------------------
def unet_forward(sd: Mapping[str, Tensor], x: Tensor, prefix: str = "encoder.unet.", capture: Optional[dict] = None) -> Tensor:
    """encoder/unet.py:225-239 with DownConv :66-72 and UpConv :101-114, built as
    UNet(32, in_channels=32, depth=3, start_filts=32, merge_mode='concat').
    `capture` (tests) receives every intermediate activation under the names the CUDA
    library uses (d0c1, d0c2, p0, ..., u1c2)."""
    g = lambda n: sd[prefix + n]
    cap = (lambda k, v: capture.__setitem__(k, v)) if capture is not None else (lambda k, v: None)
    enc = []
    depth = 3
    for i in range(depth):
        x = F.relu(F.conv2d(x, g(f"down_convs.{i}.conv1.weight"), g(f"down_convs.{i}.conv1.bias"), padding=1))
        cap(f"d{i}c1", x)
        x = F.relu(F.conv2d(x, g(f"down_convs.{i}.conv2.weight"), g(f"down_convs.{i}.conv2.bias"), padding=1))
        cap(f"d{i}c2", x)
        enc.append(x)
        if i < depth - 1:
            x = F.max_pool2d(x, kernel_size=2, stride=2)
            cap(f"p{i}", x)
    for i in range(depth - 1):
        skip = enc[-(i + 2)]
        up = F.conv_transpose2d(x, g(f"up_convs.{i}.upconv.weight"), g(f"up_convs.{i}.upconv.bias"), stride=2)
        cap(f"u{i}", up)
        x = torch.cat((up, skip), 1)
        x = F.relu(F.conv2d(x, g(f"up_convs.{i}.conv1.weight"), g(f"up_convs.{i}.conv1.bias"), padding=1))
        cap(f"u{i}c1", x)
        x = F.relu(F.conv2d(x, g(f"up_convs.{i}.conv2.weight"), g(f"up_convs.{i}.conv2.bias"), padding=1))
        cap(f"u{i}c2", x)
    return F.conv2d(x, g("conv_final.weight"), g("conv_final.bias"))


# --------------------------------------------------------------------------
# encoder/voxels.py
# --------------------------------------------------------------------------
def voxel_features(sd: Mapping[str, Tensor], x: Tensor) -> Tuple[Tensor, Tensor]:
    """encoder/voxels.py:95-108 -- voxel-centre coordinates p (B,V,3) and
    c = relu(conv_in(x)) flattened to (B,V,32)."""
    B = x.size(0)
    n_voxel = x.size(1) * x.size(2) * x.size(3)
    c1 = torch.linspace(-0.5, 0.5, x.size(1)).view(1, -1, 1, 1).expand_as(x)
    c2 = torch.linspace(-0.5, 0.5, x.size(2)).view(1, 1, -1, 1).expand_as(x)
    c3 = torch.linspace(-0.5, 0.5, x.size(3)).view(1, 1, 1, -1).expand_as(x)
    p = torch.stack([c1, c2, c3], dim=4).view(B, n_voxel, -1)
    c = F.relu(F.conv3d(x.unsqueeze(1), sd["encoder.conv_in.weight"], sd["encoder.conv_in.bias"], padding=1))
    c = c.view(B, c.size(1), -1).permute(0, 2, 1)
    return p, c


def plane_features_pre_unet(sd: Mapping[str, Tensor], x: Tensor, reso: int = 40) -> Dict[str, Tensor]:
    """encoder/voxels.py:57-66 for each plane, stopping before the U-Net."""
    p, c = voxel_features(sd, x)
    out = {}
    for plane in PLANES:
        xy = normalize_coordinate(p.clone(), plane=plane, padding=0.0)
        index = coordinate2index(xy, reso)
        fea = c.new_zeros(p.size(0), c.size(2), reso ** 2)
        fea = scatter_mean(c.permute(0, 2, 1), index, out=fea)
        out[plane] = fea.reshape(p.size(0), c.size(2), reso, reso)
    return out


def encode_inputs(sd: Mapping[str, Tensor], x: Tensor) -> Dict[str, Tensor]:
    """LocalVoxelEncoder.forward, encoder/voxels.py:89-121 -> {'xz','xy','yz'} of (B,32,40,40)."""
    pre = plane_features_pre_unet(sd, x)
    return {k: unet_forward(sd, v) for k, v in pre.items()}


# --------------------------------------------------------------------------
# conv_onet/models/decoder.py + layers.py
# --------------------------------------------------------------------------
def sample_plane_feature(p: Tensor, c: Tensor, plane: str) -> Tensor:
    """decoder.py:117-122 -- bilinear grid_sample, border padding, align_corners=True."""
    xy = normalize_coordinate(p.clone(), plane=plane, padding=0.0)
    xy = xy[:, :, None].float()
    vgrid = 2.0 * xy - 1.0
    return F.grid_sample(c, vgrid, padding_mode="border", align_corners=True, mode="bilinear").squeeze(-1)


def sample_concat_feature(p: Tensor, planes: Mapping[str, Tensor]) -> Tensor:
    """decoder.py:141-147 (concat_feat=True) -> (B,N,96), order xz|xy|yz."""
    c = [sample_plane_feature(p, planes[k], k) for k in PLANES]
    return torch.cat(c, dim=1).transpose(1, 2)


def query_feature(p: Tensor, planes: Mapping[str, Tensor]) -> Tensor:
    """decoder.py:178-191 -- the SUM (not concat) variant -> (B,N,32)."""
    c = 0
    for k in PLANES:
        c = c + sample_plane_feature(p, planes[k], k)
    return c.transpose(1, 2)


def resnet_block_fc(sd: Mapping[str, Tensor], prefix: str, x: Tensor) -> Tensor:
    """layers.py:39-47 (size_in == size_out -> identity shortcut)."""
    net = F.linear(F.relu(x), sd[prefix + "fc_0.weight"], sd[prefix + "fc_0.bias"])
    dx = F.linear(F.relu(net), sd[prefix + "fc_1.weight"], sd[prefix + "fc_1.bias"])
    return x + dx


def compute_out(sd: Mapping[str, Tensor], head: str, p: Tensor, c: Tensor) -> Tensor:
    """decoder.py:160-176 / 193-206 -- fc_p, 5x(fc_c + ResnetBlockFC), fc_out, squeeze(-1)."""
    pre = f"decoder_{head}."
    net = F.linear(p.float(), sd[pre + "fc_p.weight"], sd[pre + "fc_p.bias"])
    for i in range(5):
        net = net + F.linear(c, sd[pre + f"fc_c.{i}.weight"], sd[pre + f"fc_c.{i}.bias"])
        net = resnet_block_fc(sd, pre + f"blocks.{i}.", net)
    out = F.linear(F.relu(net), sd[pre + "fc_out.weight"], sd[pre + "fc_out.bias"])
    return out.squeeze(-1)


def local_decoder(sd: Mapping[str, Tensor], head: str, p: Tensor, planes: Mapping[str, Tensor]) -> Tensor:
    """LocalDecoder.forward, decoder.py:133-176."""
    return compute_out(sd, head, p, sample_concat_feature(p, planes))


# --------------------------------------------------------------------------
# conv_onet/models/__init__.py
# --------------------------------------------------------------------------
def decode(sd: Mapping[str, Tensor], p: Tensor, planes: Mapping[str, Tensor]):
    """models/__init__.py:111-124 -- sigmoid(qual), L2-normalised rot, raw width."""
    qual = torch.sigmoid(local_decoder(sd, "qual", p, planes))
    rot = F.normalize(local_decoder(sd, "rot", p, planes), dim=2)
    width = local_decoder(sd, "width", p, planes)
    return qual, rot, width


def forward(sd: Mapping[str, Tensor], x: Tensor, p: Tensor, p_tsdf: Optional[Tensor] = None, detach_tsdf: bool = False):
    """ConvolutionalOccupancyNetwork.forward, models/__init__.py:42-67 (detach_tsdf: :61-63, the `giga_detach` variant)."""
    planes = encode_inputs(sd, x)
    qual, rot, width = decode(sd, p, planes)
    if p_tsdf is None:
        return qual, rot, width
    if detach_tsdf:
        planes = {k: v.detach() for k, v in planes.items()}
    tsdf = local_decoder(sd, "tsdf", p_tsdf, planes)
    return qual, rot, width, tsdf


def infer_geo(sd: Mapping[str, Tensor], x: Tensor, p_tsdf: Tensor) -> Tensor:
    """models/__init__.py:69-72."""
    return local_decoder(sd, "tsdf", p_tsdf, encode_inputs(sd, x))


# --------------------------------------------------------------------------
# seeded synthetic parameters / inputs shared by the golden script, tests, bench
# --------------------------------------------------------------------------
def param_shapes(with_tsdf: bool = True, grasp_heads: bool = True) -> "list[tuple[str, tuple]]":
    """The reference's state_dict key order and shapes (probed)."""
    s = [("encoder.conv_in.weight", (32, 1, 3, 3, 3)), ("encoder.conv_in.bias", (32,))]
    for i, (ci, co) in enumerate([(32, 32), (32, 64), (64, 128)]):
        s += [(f"encoder.unet.down_convs.{i}.conv1.weight", (co, ci, 3, 3)), (f"encoder.unet.down_convs.{i}.conv1.bias", (co,)),
              (f"encoder.unet.down_convs.{i}.conv2.weight", (co, co, 3, 3)), (f"encoder.unet.down_convs.{i}.conv2.bias", (co,))]
    for i, (ci, co) in enumerate([(128, 64), (64, 32)]):
        s += [(f"encoder.unet.up_convs.{i}.upconv.weight", (ci, co, 2, 2)), (f"encoder.unet.up_convs.{i}.upconv.bias", (co,)),
              (f"encoder.unet.up_convs.{i}.conv1.weight", (co, 2 * co, 3, 3)), (f"encoder.unet.up_convs.{i}.conv1.bias", (co,)),
              (f"encoder.unet.up_convs.{i}.conv2.weight", (co, co, 3, 3)), (f"encoder.unet.up_convs.{i}.conv2.bias", (co,))]
    s += [("encoder.unet.conv_final.weight", (32, 32, 1, 1)), ("encoder.unet.conv_final.bias", (32,))]
    heads = (["qual", "rot", "width"] if grasp_heads else []) + (["tsdf"] if with_tsdf else [])
    for h in heads:
        od = 4 if h == "rot" else 1
        pre = f"decoder_{h}."
        for i in range(5):
            s += [(pre + f"fc_c.{i}.weight", (32, 96)), (pre + f"fc_c.{i}.bias", (32,))]
        s += [(pre + "fc_p.weight", (32, 3)), (pre + "fc_p.bias", (32,))]
        for i in range(5):
            s += [(pre + f"blocks.{i}.fc_0.weight", (32, 32)), (pre + f"blocks.{i}.fc_0.bias", (32,)),
                  (pre + f"blocks.{i}.fc_1.weight", (32, 32)), (pre + f"blocks.{i}.fc_1.bias", (32,))]
        s += [(pre + "fc_out.weight", (od, 32)), (pre + "fc_out.bias", (od,))]
    return s


def seeded_state_dict(seed: int = 1, std: float = 0.1, **kw) -> Dict[str, Tensor]:
    """Every parameter ~ N(0, std^2) from numpy's frozen legacy MT19937 stream, drawn in
    the reference state_dict order (default init zeroes fc_1.weight and U-Net biases and
    would hide errors -- SURVEY.md section 7 step 0)."""
    import numpy as np
    rs = np.random.RandomState(seed)
    return {k: torch.from_numpy((rs.standard_normal(shp) * std).astype(np.float32)) for k, shp in param_shapes(**kw)}


def seeded_inputs(B: int, N: int, seed: int = 0, edge_cases: bool = True):
    """TSDF in [0,1), grasp / occupancy query points in [-0.5,0.5); with
    `edge_cases` a few points are pushed onto +-0.5 and outside the cube to hit both
    clamps of normalize_coordinate (occupancy points may lie outside, utils/implicit.py:78-84)."""
    import numpy as np
    rs = np.random.RandomState(1000 + seed)
    x = rs.random_sample((B, 40, 40, 40)).astype(np.float32)
    p = (rs.random_sample((B, N, 3)) - 0.5).astype(np.float32)
    pt = (rs.random_sample((B, N, 3)) - 0.5).astype(np.float32)
    if edge_cases and N >= 8:
        for a in (p, pt):
            a[:, 0] = 0.5
            a[:, 1] = -0.5
            a[:, 2] = (0.5, -0.5, 0.0)
            a[:, 3] = (0.62, -0.57, 0.51)
            a[:, 4] = (-0.5000001, 0.4999999, 0.5000001)
            a[:, 5] = 0.0
            a[:, 6] = (1.0 / 39 - 0.5, 2.0 / 39 - 0.5, 38.0 / 39 - 0.5)
    return torch.from_numpy(x), torch.from_numpy(p), torch.from_numpy(pt)
