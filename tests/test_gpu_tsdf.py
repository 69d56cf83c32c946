"""GPU: TSDF integration / read-out kernels (csrc/tsdf.cuh, giga_tsdf_integrate / giga_tsdf_grid) against the numpy restatement of
Open3D's UniformTSDFVolume (oracle/tsdf_oracle.py), bit for bit: the volume, the weights and the network input grid; view order, 40^3
and 120^3 volumes, more than 16 views (two launches), and perception -> network without leaving the device."""
import numpy as np
import pytest
import torch

from giga_b200 import perception
from oracle import tsdf_oracle as TO

pytestmark = pytest.mark.gpu


def _intr(i):
    return perception.CameraIntrinsic(i.width, i.height, i.fx, i.fy, i.cx, i.cy)


@pytest.mark.parametrize("seed,R,n_views", [(1, 40, 6), (2, 40, 1), (4, 120, 3), (5, 40, 20)])
def test_integration_and_grid_bit_exact(seed, R, n_views):
    imgs, intr, Ts = TO.seeded_scene(seed, n_views=n_views)
    ref = TO.create_tsdf(0.3, R, imgs, intr, Ts)
    vol = perception.create_tsdf(0.3, R, imgs, _intr(intr), Ts)
    assert np.array_equal(vol._weight.cpu().numpy(), ref.weight)
    assert np.array_equal(vol._tsdf.cpu().numpy(), ref.tsdf)
    g = vol.get_grid()
    assert g.shape == (1, R, R, R) and g.dtype == np.float32
    assert np.array_equal(g, ref.get_grid())
    assert ref.weight.max() == n_views


def test_other_camera_and_volume_geometry():
    """a non-square image with an off-centre principal point, a larger volume, views from below the table plane (many voxels behind or
    outside the frustum), and a depth image with holes (zeros) and far readings (>= depth_trunc)"""
    rng = np.random.default_rng(3)
    intr = TO.Intrinsic(200, 150, 180.0, 175.0, 90.5, 80.25)
    boxes = [((0.1, 0.1, 0.05), (0.2, 0.25, 0.3)), ((0.3, 0.3, 0.05), (0.38, 0.36, 0.12))]
    Ts, imgs = [], []
    for k, eye in enumerate([(0.9, 0.2, 0.6), (-0.4, 0.25, 0.5), (0.25, 0.25, 1.2), (0.25, -0.6, -0.1)]):
        T = TO.look_at(eye, (0.25, 0.25, 0.1))
        d = TO.render_depth(intr, T, boxes, noise=0.001, rng=rng)
        d[rng.random(d.shape) < 0.05] = 0.0           # holes
        d[:10] = 2.5                                    # beyond depth_trunc: dropped at RGBD creation
        Ts.append(T); imgs.append(d)
    imgs, Ts = np.stack(imgs), np.stack(Ts)
    for R in (40, 64):
        ref = TO.create_tsdf(0.5, R, imgs, intr, Ts)
        vol = perception.create_tsdf(0.5, R, imgs, _intr(intr), Ts)
        assert np.array_equal(vol._weight.cpu().numpy(), ref.weight) and np.array_equal(vol._tsdf.cpu().numpy(), ref.tsdf)
        assert np.array_equal(vol.get_grid(), ref.get_grid())
        assert 0 < (ref.weight > 0).mean() < 1


def test_incremental_integrate_equals_batched_and_order_matters():
    imgs, intr, Ts = TO.seeded_scene(7, n_views=5)
    a = perception.TSDFVolume(0.3, 40)
    for k in range(5):
        a.integrate(imgs[k], _intr(intr), Ts[k])          # the reference's call pattern (simulation.py:175-180)
    b = perception.create_tsdf(0.3, 40, imgs, _intr(intr), Ts)
    assert torch.equal(a._tsdf, b._tsdf) and torch.equal(a._weight, b._weight)
    ref_rev = TO.create_tsdf(0.3, 40, imgs[::-1], intr, Ts[::-1])
    c = perception.create_tsdf(0.3, 40, imgs[::-1].copy(), _intr(intr), Ts[::-1].copy())
    assert np.array_equal(c._tsdf.cpu().numpy(), ref_rev.tsdf)


def test_quaternion_extrinsics_and_device_grid_feeds_the_network(oracle_sd):
    from scipy.spatial.transform import Rotation
    from tests.util import make_net
    from oracle import giga_oracle as O
    imgs, intr, Ts = TO.seeded_scene(9, n_views=6)
    lists = [np.r_[Rotation.from_matrix(T[:3, :3]).as_quat(), T[:3, 3]] for T in Ts]     # Transform.to_list(), perception.py:123
    vol = perception.create_tsdf(0.3, 40, imgs, _intr(intr), lists)
    mats = []
    for l in lists:
        m = np.eye(4); m[:3, :3] = Rotation.from_quat(l[:4]).as_matrix(); m[:3, 3] = l[4:]; mats.append(m)
    ref = TO.create_tsdf(0.3, 40, imgs, intr, np.stack(mats))
    grid = vol.get_grid_device()
    assert grid.is_cuda and np.array_equal(grid.cpu().numpy(), ref.get_grid())
    net = make_net("giga", oracle_sd)
    _, p, pt = O.seeded_inputs(1, 64, seed=2)
    with torch.no_grad():
        out = net(grid, p.cuda(), p_tsdf=pt.cuda())
        want = O.forward(oracle_sd, torch.from_numpy(ref.get_grid()), p, pt)
    for a, b in zip(out, want):
        assert (a.cpu() - b).abs().max().item() <= 1e-4
    with pytest.raises(Exception):
        vol.integrate(imgs[0][:100], _intr(intr), Ts[0])
