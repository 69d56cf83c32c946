"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the
header declares, the host mirror has the reference's parameter inventory, and the product never
touches the oracle or computes on the CPU."""
import ctypes as C
import os
import re

import pytest
import torch

from tests.util import ROOT


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "giga_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(giga_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import giga_b200
    from giga_b200 import _lib

    declared = _declared_symbols()
    assert len(declared) >= 14
    raw = C.CDLL(_lib.LIB_PATH)
    for s in declared:
        assert hasattr(raw, s), f"{s} declared in include/giga_b200.h but not exported"
        assert s in _lib.SYMBOLS, f"{s} has no ctypes prototype"
    assert sorted(_lib.SYMBOLS) == declared
    assert raw.giga_version() == 100
    assert os.path.dirname(_lib.LIB_PATH) == os.path.join(ROOT, "giga_b200")  # in-tree, not site-packages


def test_state_dict_matches_reference_inventory(golden):
    import giga_b200

    for name in ["giga", "giga_aff", "giga_geo", "giga_detach"]:
        net = giga_b200.get_network(name)
        sd = net.state_dict()
        assert list(sd.keys()) == list(golden[f"keys_{name}"])
        assert [str(tuple(v.shape)) for v in sd.values()] == list(golden[f"shapes_{name}"])
    assert sum(p.numel() for p in giga_b200.get_network("giga").parameters()) == 581863
    assert sum(p.numel() for p in giga_b200.get_network("vgn").parameters()) == 313238      # the baseline ConvNet (networks.py:48-63)
    with pytest.raises(KeyError):
        giga_b200.get_network("nope")


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    import giga_b200
    from giga_b200._lib import lib

    net = giga_b200.get_network("giga")
    with pytest.raises(giga_b200.GigaError):
        net.encode_inputs(torch.zeros(1, 40, 40, 40))
    with pytest.raises(giga_b200.GigaError):
        net(torch.zeros(1, 40, 40, 40), torch.zeros(1, 4, 3))
    h = C.c_void_p()
    assert lib.giga_ctx_create(C.byref(h), 0) == -2  # GIGA_ENODEV
    assert b"no CPU path" in lib.giga_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "giga_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"(import\s+oracle|from\s+oracle|oracle/|oracle\.)", txt), f"{f} uses the oracle"
    for f in ["bench.py", "__graft_entry__.py"]:
        p = os.path.join(ROOT, f)
        if os.path.exists(p):
            assert "/root/reference" not in open(p).read(), f"{f} must not read /root/reference at run time"


def test_planner_host_mirror_surface():
    """The planner mirror has the reference's constructor / call surface (detection_implicit.py:17-33) and the parameter struct
    carries the reference's defaults (process/bound/select); no GPU needed for any of this."""
    import inspect

    import giga_b200
    from giga_b200._lib import SelectParams, lib
    from giga_b200.detection_implicit import LOW_TH, VGNImplicit, lattice, select_params

    sig = inspect.signature(VGNImplicit.__init__)
    assert list(sig.parameters)[:9] == ["self", "model_path", "model_type", "best", "force_detection", "qual_th", "out_th", "visualize", "resolution"]
    assert (sig.parameters["qual_th"].default, sig.parameters["out_th"].default, sig.parameters["resolution"].default) == (0.9, 0.5, 40)
    assert list(inspect.signature(VGNImplicit.__call__).parameters) == ["self", "state", "scene_mesh", "aff_kwargs"]
    p = SelectParams()
    lib.giga_select_params_default(C.byref(p))
    q = select_params()
    for f, _ in SelectParams._fields_:
        assert getattr(p, f) == getattr(q, f), f
    assert (p.lim_x, p.lim_y, p.lim_z, p.max_filter_size, p.force_detection) == (2, 2, 7, 4, 0) and LOW_TH == 0.5
    pos = lattice()
    assert pos.shape == (1, 64000, 3) and pos.dtype == torch.float32
    lin = torch.linspace(start=-0.5, end=0.5 - 1.0 / 40, steps=40)
    assert torch.equal(pos[0, :40, 2], lin) and torch.equal(pos[0, ::1600, 0], lin)     # meshgrid 'ij': z fastest, x slowest
    if not torch.cuda.is_available():
        with pytest.raises(giga_b200.GigaError):
            VGNImplicit(None, "giga")
        # compute entry points refuse to run without a device (no context can exist)
        assert lib.giga_select_grasps(None, None, None, None, None, 1, C.byref(p), 8, None, None, None, None, None, None, None) == -1
        assert lib.giga_detect_host(None, None, None, 1, C.byref(p), 8, None, None, None, None, None, None) == -1
        assert lib.giga_forward(None, None, 1, None, 0, None, 0, None, None, None, None, None, None, None, None) == -1


def test_training_and_perception_host_surface():
    """The training step and the perception mirror keep the reference's surface (train_giga.py:67,153-174; perception.py:10-126) and
    refuse to compute without a device: no PyTorch / CPU path stands in for the kernels."""
    import inspect

    import numpy as np

    import giga_b200
    from giga_b200 import perception, training
    from giga_b200._lib import lib

    assert list(inspect.signature(training.loss_fn).parameters) == ["y_pred", "y"]
    assert list(inspect.signature(training.select).parameters) == ["out"]
    assert list(inspect.signature(training.Adam.__init__).parameters)[:6] == ["self", "params", "lr", "betas", "eps", "weight_decay"]
    assert list(inspect.signature(perception.TSDFVolume.__init__).parameters)[:3] == ["self", "size", "resolution"]
    assert list(inspect.signature(perception.TSDFVolume.integrate).parameters) == ["self", "depth_img", "intrinsic", "extrinsic"]
    assert list(inspect.signature(perception.create_tsdf).parameters) == ["size", "resolution", "depth_imgs", "intrinsic", "extrinsics"]
    intr = perception.CameraIntrinsic(640, 480, 540.0, 541.0, 320.0, 240.0)
    assert (intr.fx, intr.fy, intr.cx, intr.cy) == (540.0, 541.0, 320.0, 240.0)
    assert perception.CameraIntrinsic.from_dict(intr.to_dict()).K.tolist() == intr.K.tolist()
    if not torch.cuda.is_available():
        net = giga_b200.get_network("giga")                         # trainable parameters + gradient mode: the training path
        assert net._train_active(torch.zeros(1, 4, 3))
        with pytest.raises(giga_b200.GigaError):
            net(torch.zeros(1, 40, 40, 40), torch.zeros(1, 1, 3), p_tsdf=torch.zeros(1, 8, 3))
        with pytest.raises(giga_b200.GigaError):
            perception.TSDFVolume(0.3, 40)
        assert lib.giga_train_forward(None, None, 1, None, 0, None, 0, 0, None, None, None, None, None) == -1
        assert lib.giga_train_backward(None, None, None, None, None, None, None) == -1
        assert lib.giga_tsdf_grid(None, None, None, 40, None, None) == -1
