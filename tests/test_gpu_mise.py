"""GPU parity of the device-side Generator3D occupancy sweep (giga_mise_sweep): the MISE bookkeeping is bit-exact against the CPU oracle
(itself pinned bit for bit against the compiled reference octree) when both see the same network values, and the value grid matches the
golden fixture produced by the UNMODIFIED reference pipeline (reference network + Generator3D loop + reference MISE)."""
import os

import numpy as np
import pytest
import torch

from oracle import giga_oracle as O
from oracle import mise_oracle as M
from tests.util import make_net

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def net(oracle_sd):
    return make_net("giga", oracle_sd)


@pytest.mark.parametrize("r0,steps,th,pad", [(8, 2, 0.5, 0.0), (16, 3, 0.5, 0.1), (4, 3, 0.45, 0.1), (5, 1, 0.6, 0.0)])
def test_device_sweep_equals_oracle_loop_on_gpu_values(net, r0, steps, th, pad):
    """Same values in, same octree out: the oracle MISE is driven by the GPU's own decode_occ outputs (the reference's host loop), the
    device sweep does everything in one call.  Grid, iteration count and number of evaluated points must be identical."""
    from giga_b200.generation import Generator3D
    x, _, _ = O.seeded_inputs(1, 8, seed=21)
    with torch.no_grad():
        c = net.encode_inputs(x.to(DEV))
        gen = Generator3D(net, device=torch.device(DEV), threshold=th, input_type="pointcloud", padding=pad, resolution0=r0, upsampling_steps=steps)
        stats = {}
        grid = gen.generate_value_grid(c, stats).cpu().numpy()

        def eval_points(pf):
            return net.decode_occ(torch.from_numpy(pf).to(DEV)[None], c).logits[0].cpu().numpy()

        ref_grid, iters, total, _ = M.sweep(eval_points, r0, steps, th, pad)
    assert stats["mise iterations"] == iters and stats["points evaluated"] == total, (stats, iters, total)
    assert np.array_equal(grid, ref_grid.astype(np.float32))
    assert total < ((r0 << steps) + 1) ** 3          # the sweep refined only near the surface


def test_value_grid_matches_reference_pipeline_golden(net):
    from giga_b200.generation import Generator3D
    g = np.load(os.path.join(ROOT, "tests", "golden", "mise_golden.npz"))
    x = torch.from_numpy(g["net_x"])
    with torch.no_grad():
        c = net.encode_inputs(x.to(DEV))
        for key in ("c", "d"):
            r0, steps, th, pad = g[key + "_cfg"]
            assert float(g[key + "_margin"]) > 2e-4          # no evaluated point lies within the network tolerance of the threshold
            gen = Generator3D(net, device=torch.device(DEV), threshold=float(th), input_type="pointcloud", padding=float(pad),
                              resolution0=int(r0), upsampling_steps=int(steps))
            grid = gen.generate_value_grid(c).cpu().numpy()
            assert grid.shape == g[key + "_grid"].shape
            assert np.abs(grid - g[key + "_grid"]).max() <= 1e-4, key    # same refinement pattern, values within the network tolerance


def test_regular_grid_shortcut_and_errors(net):
    from giga_b200.generation import Generator3D
    x, _, _ = O.seeded_inputs(2, 8, seed=22)
    with torch.no_grad():
        c1 = net.encode_inputs(x[:1].to(DEV))
        gen0 = Generator3D(net, device=torch.device(DEV), padding=0.0, resolution0=12, upsampling_steps=0)
        grid = gen0.generate_value_grid(c1)
        lin = torch.linspace(-0.5, 0.5, 12)
        p = torch.stack(torch.meshgrid(lin, lin, lin, indexing="ij"), -1).reshape(1, -1, 3)
        ref = O.local_decoder(O.seeded_state_dict(seed=1), "tsdf", p, O.encode_inputs(O.seeded_state_dict(seed=1), x[:1]))
        assert (grid.cpu().reshape(-1) - ref.reshape(-1)).abs().max().item() <= 1e-4
        c2 = net.encode_inputs(x.to(DEV))
        with pytest.raises(Exception):
            Generator3D(net, device=torch.device(DEV)).generate_value_grid(c2)          # one scene at a time
        with pytest.raises(NotImplementedError):
            Generator3D(net, device=torch.device(DEV), resolution0=4, upsampling_steps=1).generate_from_latent(c1)   # no libmcubes / trimesh here
