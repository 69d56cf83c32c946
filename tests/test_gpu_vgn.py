"""GPU: the VGN baseline network (csrc/vgn.cuh behind giga_vgn_forward) against the oracle and the reference-made fixture, and the VGN
planner mirror (giga_b200.detection.VGN) against the numpy planner oracle with the VGN width gate."""
import os

import numpy as np
import pytest
import torch

import giga_b200
from oracle import planner_oracle as P
from oracle import vgn_oracle as V

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "vgn_golden.npz")
TOL = 1e-4   # north_star tolerance for fp32 outputs


def _net(sd):
    net = giga_b200.get_network("vgn")
    net.load_state_dict(sd)
    return net.to("cuda").eval()


def test_vgn_forward_matches_oracle_and_reference_fixture():
    gold = np.load(GOLD)
    sd = V.seeded_state_dict(seed=3)
    x = V.seeded_inputs(2, seed=5)
    net = _net(sd)
    with torch.no_grad():
        qual, rot, width = net(x.cuda())
    assert qual.shape == (2, 1, 40, 40, 40) and rot.shape == (2, 4, 40, 40, 40) and width.shape == (2, 1, 40, 40, 40)
    ref = V.forward(sd, x)
    vox = gold["vox"]
    for name, a, b in zip(("qual", "rot", "width"), (qual, rot, width), ref):
        a = a.cpu()
        err = (a - b).abs().max().item()
        assert err < TOL, f"{name}: {err}"
        got = a.reshape(a.shape[0], a.shape[1], 64000)[:, :, vox].numpy()
        np.testing.assert_allclose(got, gold[name], rtol=0, atol=TOL, err_msg=name)
    assert net.gpu_launches == 7


@pytest.mark.parametrize("B", [1, 5])
def test_vgn_batches_and_default_init(B):
    torch.manual_seed(B)
    net = giga_b200.get_network("vgn")       # nn.Conv3d default initialisation
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    net = net.to("cuda")
    x = V.seeded_inputs(B, seed=B)
    with torch.no_grad():
        out = net(x.cuda())
        out1 = net(x[B - 1:].cuda())         # the last scene alone: batch independence, workspace reuse at a smaller B
    ref = V.forward(sd, x)
    for name, a, b, c in zip(("qual", "rot", "width"), out, ref, out1):
        assert (a.cpu() - b).abs().max().item() < TOL, name
        assert torch.equal(a[B - 1:], c), name
    # zero rotation logits: F.normalize's eps path (0 / max(0, 1e-12) = 0)
    with torch.no_grad():
        net.conv_rot.weight.zero_()
        net.conv_rot.bias.zero_()
        _, rot, _ = net(x.cuda())
    assert float(rot.abs().max()) == 0.0


def test_vgn_rejects_bad_input():
    net = _net(V.seeded_state_dict(seed=3))
    with pytest.raises(giga_b200.GigaError):
        net(torch.zeros(1, 1, 32, 32, 32, device="cuda"))
    giga = giga_b200.get_network("giga").to("cuda")
    with pytest.raises(giga_b200.GigaError):      # a GIGA context has no VGN parameters
        from giga_b200.networks import ConvNet
        ConvNet.forward_flat(giga, torch.zeros(1, 1, 40, 40, 40, device="cuda"))


class _State:
    def __init__(self, tsdf):
        self.tsdf = tsdf


@pytest.mark.parametrize("force", [False, True])
def test_vgn_planner_matches_reference_pipeline(force):
    """VGN.__call__ (detection.py:37-81): predict -> process (width gate 1.33..9.33 voxels) -> bound -> select -> from_voxel_coordinates."""
    sd = V.seeded_state_dict(seed=3)
    x = V.seeded_inputs(1, seed=5 if not force else 6)
    qual_th = 0.9 if not force else 0.9999999
    planner = giga_b200.VGN(None, "vgn", best=True, qual_th=qual_th, force_detection=force)
    planner.net = _net(sd)
    tsdf = x[0].numpy()                                   # (1,40,40,40)
    grasps, scores, toc = planner(_State(tsdf))
    # expected: the planner oracle on the DEVICE network outputs (the network itself is checked above; thresholds are discontinuous)
    with torch.no_grad():
        q, r, w = planner.net.forward_flat(x.cuda())
    idx, sc, rot, width, _ = P.detect(tsdf, q[0].cpu().numpy(), r[0].cpu().numpy(), w[0].cpu().numpy(), qual_th=qual_th,
                                      force_detection=force, min_width=1.33, max_width=9.33)
    assert len(grasps) == len(idx) and len(idx) > 0
    if force:
        assert len(grasps) == 1
    vs = 0.3 / 40
    np.testing.assert_array_equal(np.asarray(scores, np.float32), sc)
    for g, i, ro, wi in zip(grasps, idx, rot, width):
        np.testing.assert_array_equal(g.pose.translation, i.astype(np.float64) * vs)
        np.testing.assert_array_equal(g.pose.rotation.as_quat(), giga_b200.detection_implicit.Rotation.from_quat(ro).as_quat())
        assert g.width == wi * vs
    assert toc > 0
