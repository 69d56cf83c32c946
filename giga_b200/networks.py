"""Model factory with the reference's entry points (/root/reference/src/vgn/networks.py:10-35).

`get_network(name)` / `load_network(path, device, model_type)` return objects with the reference
model interface, backed by the sm_100a CUDA library.  `vgn` is the dense 3-D ConvNet baseline
(networks.py:48-63): `ConvNet` below is its parameter container, the
arithmetic is giga_vgn_forward (csrc/vgn.cuh).
"""
from __future__ import annotations

import torch

import ctypes as C
import math

import torch.nn as nn

from . import _lib
from ._lib import check, lib
from .model import ConvolutionalOccupancyNetwork, ConvolutionalOccupancyNetworkGeometry, _Engine, _prep, _stream


class _Conv3dParams(nn.Module):
    """weight/bias holder with nn.Conv3d's names, shapes and default initialisation (kaiming-uniform a=sqrt(5))."""

    def __init__(self, c_in, c_out, k):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(c_out, c_in, k, k, k))
        self.bias = nn.Parameter(torch.empty(c_out))
        bound = 1.0 / math.sqrt(c_in * k ** 3)
        nn.init.uniform_(self.weight, -bound, bound)
        nn.init.uniform_(self.bias, -bound, bound)


class _Stack(nn.Module):
    def __init__(self, c_in, filters, kernels):
        super().__init__()
        self.conv1 = _Conv3dParams(c_in, filters[0], kernels[0])
        self.conv2 = _Conv3dParams(filters[0], filters[1], kernels[1])
        self.conv3 = _Conv3dParams(filters[1], filters[2], kernels[2])


class ConvNet(nn.Module):
    """Drop-in for networks.py:48-63 (`get_network("vgn")`): same state_dict keys / shapes, forward(x) with x (B,1,40,40,40) ->
    qual (B,1,40,40,40), rot (B,4,40,40,40), width (B,1,40,40,40).  Parameter container only; no CPU path."""

    def __init__(self):
        super().__init__()
        self.encoder = _Stack(1, [16, 32, 64], [5, 3, 3])
        self.decoder = _Stack(64, [64, 32, 16], [3, 3, 5])
        self.conv_qual = _Conv3dParams(16, 1, 5)
        self.conv_rot = _Conv3dParams(16, 4, 5)
        self.conv_width = _Conv3dParams(16, 1, 5)

    def _engine(self) -> _Engine:
        dev = next(self.parameters()).device
        eng = self.__dict__.get("_eng")
        if eng is None or eng.device != dev:
            eng = _Engine(dev)
            self.__dict__["_eng"] = eng
        eng.sync_params(self)
        return eng

    def invalidate_params(self):
        eng = self.__dict__.get("_eng")
        if eng is not None:
            eng.invalidate()
        return self

    def __getstate__(self):
        d = self.__dict__.copy()
        d.pop("_eng", None)
        return d

    @property
    def gpu_launches(self) -> int:
        eng = self.__dict__.get("_eng")
        return eng.launches if eng else 0

    def forward_flat(self, x: torch.Tensor):
        """-> qual (B,64000), rot (B,64000,4), width (B,64000): the layouts giga_select_grasps takes."""
        eng = self._engine()
        x = _prep(x, eng.device)
        if x.dim() == 5 and x.shape[1] == 1:
            x = x[:, 0]
        if x.dim() != 4 or tuple(x.shape[1:]) != (40, 40, 40):
            raise _lib.GigaError(f"inputs must be (B,1,40,40,40), got {tuple(x.shape)}")
        x = x.contiguous()
        B = x.shape[0]
        mk = lambda *s: torch.empty(s, device=eng.device, dtype=torch.float32)
        qual, rot, width = mk(B, 64000), mk(B, 64000, 4), mk(B, 64000)
        check(lib.giga_vgn_forward(eng.h, C.c_void_p(x.data_ptr()), B, C.c_void_p(qual.data_ptr()), C.c_void_p(rot.data_ptr()),
                                   C.c_void_p(width.data_ptr()), _stream(eng.device)), "giga_vgn_forward")
        return qual, rot, width

    def forward(self, x):
        """networks.py:57-63"""
        qual, rot, width = self.forward_flat(x)
        B = qual.shape[0]
        return qual.view(B, 1, 40, 40, 40), rot.view(B, 40, 40, 40, 4).permute(0, 4, 1, 2, 3), width.view(B, 1, 40, 40, 40)


def GIGAAff():      # networks.py:65-89  (no decoder_tsdf)
    return ConvolutionalOccupancyNetwork(with_tsdf=False)


def GIGA():         # networks.py:91-115
    return ConvolutionalOccupancyNetwork(with_tsdf=True)


def GIGAGeo():      # networks.py:117-142 (tsdf_only)
    return ConvolutionalOccupancyNetworkGeometry()


def GIGADetach():   # networks.py:144-169
    return ConvolutionalOccupancyNetwork(with_tsdf=True, detach_tsdf=True)


def get_network(name):
    models = {
        "giga_aff": GIGAAff,
        "giga": GIGA,
        "giga_geo": GIGAGeo,
        "giga_detach": GIGADetach,
    }
    models["vgn"] = ConvNet
    return models[name.lower()]()


def load_network(path, device, model_type=None):
    """Construct the network and load parameters from `path` (name must conform to `vgn_name_[_...]`)."""
    if model_type is None:
        model_name = "_".join(path.stem.split("_")[1:-1])
    else:
        model_name = model_type
    print(f"Loading [{model_type}] model from {path}")
    net = get_network(model_name).to(device)
    net.load_state_dict(torch.load(path, map_location=device))
    return net
