"""CPU: the VGN oracle against the fixture made from the reference ConvNet (tests/golden/make_vgn_golden.py), and the host-side
container (`get_network("vgn")`): reference state_dict names / shapes, loud failure without a GPU."""
import os

import numpy as np
import pytest
import torch

from oracle import vgn_oracle as V

GOLD = os.path.join(os.path.dirname(__file__), "golden", "vgn_golden.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_oracle_matches_reference_convnet(gold):
    sd = V.seeded_state_dict(seed=3)
    x = V.seeded_inputs(2, seed=5)
    chk = float(sum(v.double().abs().sum() for v in sd.values()) + x.double().sum())
    assert chk == float(gold["checksum"]), "seeded parameters / inputs drifted from the fixture"
    cap = {}
    qual, rot, width = V.forward(sd, x, capture=cap)
    np.testing.assert_allclose(cap["enc"].numpy(), gold["enc"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(cap["dec"][:, :, ::4, ::4, ::4].numpy(), gold["dec_s"], rtol=0, atol=2e-6)
    vox = gold["vox"]
    for name, t in (("qual", qual), ("rot", rot), ("width", width)):
        got = t.reshape(t.shape[0], t.shape[1], 64000)[:, :, vox].numpy()
        np.testing.assert_allclose(got, gold[name], rtol=0, atol=5e-6, err_msg=name)
    sums = np.array([qual.double().sum(), rot.double().abs().sum(), width.double().sum()])
    np.testing.assert_allclose(sums, gold["sums"], rtol=1e-6)


def test_up2_is_nearest_interpolate():
    x = torch.arange(2 * 3 * 5 ** 3, dtype=torch.float32).reshape(2, 3, 5, 5, 5)
    assert torch.equal(V._up2(x), torch.nn.functional.interpolate(x, 10))


def test_container_has_the_reference_state_dict():
    import giga_b200
    net = giga_b200.get_network("vgn")
    sd = net.state_dict()
    ref = V.seeded_state_dict(seed=0)
    assert list(sd.keys()) == list(ref.keys())          # registration order = the reference's (encoder, decoder, heads)
    for k in ref:
        assert tuple(sd[k].shape) == tuple(ref[k].shape), k
    net.load_state_dict(ref, strict=True)
    # nn.Conv3d default initialisation: U(-1/sqrt(fan_in), 1/sqrt(fan_in))
    fresh = giga_b200.get_network("vgn")
    w = fresh.encoder.conv2.weight
    bound = 1.0 / np.sqrt(16 * 27)
    assert float(w.abs().max()) <= bound and float(w.abs().max()) > 0.9 * bound


def test_no_cpu_path():
    import giga_b200
    net = giga_b200.get_network("vgn")
    with pytest.raises(giga_b200.GigaError):
        net(torch.zeros(1, 1, 40, 40, 40))
