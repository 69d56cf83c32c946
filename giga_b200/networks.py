"""Model factory with the reference's entry points (/root/reference/src/vgn/networks.py:10-35).

`get_network(name)` / `load_network(path, device, model_type)` return objects with the reference
model interface, backed by the sm_100a CUDA library.  `vgn` (the dense 3-D ConvNet baseline,
networks.py:48-63) is outside the hot path and not provided.
"""
from __future__ import annotations

import torch

from .model import ConvolutionalOccupancyNetwork, ConvolutionalOccupancyNetworkGeometry


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
    key = name.lower()
    if key == "vgn":
        raise NotImplementedError("the VGN ConvNet baseline is outside the GIGA dense-inference hot path (SURVEY.md 8f)")
    return models[key]()


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
