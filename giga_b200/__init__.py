"""giga_b200 -- B200-native (sm_100a) implementation of GIGA's dense-inference hot path.

Public surface mirrors the reference (vgn.networks / ConvONets.conv_onet.models):
    from giga_b200 import get_network, load_network
    net = get_network("giga").to("cuda")
    qual, rot, width, occ = net(tsdf, p, p_tsdf=p_occ)          # differentiable when the parameters require gradients (native backward)
also: detection_implicit.VGNImplicit (planner), perception.TSDFVolume / create_tsdf, generation.Generator3D, detection.VGN,
training.select / loss_fn / Adam.
Importing this package loads libgiga_b200.so and fails loudly if it is not built.
"""
from ._lib import GigaError, LIB_PATH, HEAD_QUAL, HEAD_ROT, HEAD_WIDTH, HEAD_TSDF, HEAD_GRASP  # noqa: F401
from .model import (ConvolutionalOccupancyNetwork, ConvolutionalOccupancyNetworkGeometry, LocalDecoder,  # noqa: F401
                    LocalVoxelEncoder, PlaneFeatures, UNet)
from .networks import get_network, load_network  # noqa: F401
from . import detection_implicit  # noqa: F401
from .detection_implicit import VGNImplicit  # noqa: F401
from . import detection  # noqa: F401
from .detection import VGN  # noqa: F401
from . import generation  # noqa: F401
from . import training  # noqa: F401
from . import perception  # noqa: F401
from .generation import Generator3D  # noqa: F401

__version__ = "0.1.0"
