"""CPU: the TSDF-integration oracle (oracle/tsdf_oracle.py; parity UNPINNED against Open3D, which is absent here) satisfies the closed-form
properties of the published algorithm: a fronto-parallel wall gives the analytic truncated ramp, unobserved / out-of-range voxels stay
empty, the running average of identical views is idempotent, and get_grid maps (-0.98, 0.98) -> (tsdf + 1) / 2."""
import numpy as np

from oracle import tsdf_oracle as TO


def _wall_setup(depth=0.4, W=64, H=48):
    intr = TO.Intrinsic(W, H, 60.0, 60.0, W / 2 - 0.5, H / 2 - 0.5)
    # camera at (0.15, 0.15, -0.25) looking along +z of the task frame: x_cam = x, y_cam = y, z_cam = z + 0.25
    T = np.eye(4)
    T[:3, 3] = [-0.15, -0.15, 0.25]
    img = np.full((H, W), depth, np.float32)
    return intr, T, img


def test_fronto_parallel_wall_matches_closed_form():
    size, R = 0.3, 40
    intr, T, img = _wall_setup()
    vol = TO.TSDFVolume(size, R)
    vol.integrate(img, intr, T)
    vl = size / R
    # on the optical axis (x = y = 0.15 -> voxel 19/20 straddle it): sdf = (d - z_cam) * multiplier, multiplier ~ 1 near the axis
    x = y = 20
    zc = vl * (np.arange(R) + 0.5) + 0.25
    px, py = (vl * 20.5 - 0.15), (vl * 20.5 - 0.15)
    u = int(px * 60.0 / zc[0] + intr.cx + 0.5)
    want_all = []
    for z in range(R):
        uu = int(np.float32(px * 60.0 / zc[z] + intr.cx + 0.5)); vv = int(np.float32(py * 60.0 / zc[z] + intr.cy + 0.5))
        mult = np.sqrt(((uu - intr.cx) / 60.0) ** 2 + ((vv - intr.cy) / 60.0) ** 2 + 1.0)
        sdf = (0.4 - zc[z]) * mult
        want_all.append(min(1.0, sdf / (4 * vl)) if sdf > -4 * vl else None)
    for z, want in enumerate(want_all):
        if want is None:
            assert vol.weight[x, y, z] == 0 and vol.tsdf[x, y, z] == 0
        else:
            assert vol.weight[x, y, z] == 1
            assert abs(vol.tsdf[x, y, z] - want) < 2e-5, (z, vol.tsdf[x, y, z], want)
    assert u >= 0
    # the wall sits at task z = 0.15: voxels in front of it are free (+), behind it negative down to the truncation, then unobserved
    assert vol.tsdf[x, y, 0] == 1.0 and vol.tsdf[x, y, 21] < 0 and vol.weight[x, y, R - 1] == 0


def test_views_average_and_trunc_and_grid():
    size, R = 0.3, 40
    intr, T, img = _wall_setup()
    one = TO.TSDFVolume(size, R); one.integrate(img, intr, T)
    three = TO.create_tsdf(size, R, np.stack([img] * 3), intr, np.stack([T] * 3))
    assert np.array_equal(three.weight, 3 * one.weight)
    np.testing.assert_allclose(three.tsdf, one.tsdf, rtol=0, atol=2e-7)          # averaging identical observations
    far = TO.TSDFVolume(size, R); far.integrate(np.full_like(img, 2.0), intr, T)   # depth >= depth_trunc is dropped (RGBD creation)
    assert far.weight.max() == 0
    g = one.get_grid()
    assert g.shape == (1, R, R, R) and g.dtype == np.float32
    keep = (one.weight != 0) & (one.tsdf < np.float32(0.98)) & (one.tsdf >= np.float32(-0.98))
    assert np.array_equal(g[0] != 0, keep | False) or np.array_equal(g[0][keep], ((one.tsdf[keep].astype(np.float64) + 1) * 0.5).astype(np.float32))
    assert g[0][~keep].max() == 0 and 0.0 < g[0][keep].min() and g[0][keep].max() < 0.99


def test_seeded_scene_is_a_plausible_scan():
    imgs, intr, Ts = TO.seeded_scene(3, n_views=6)
    vol = TO.create_tsdf(0.3, 40, imgs, intr, Ts)
    g = vol.get_grid()[0]
    occupied = (g > 0) & (g < 0.5)
    assert 0.002 < occupied.mean() < 0.5           # a table surface + a few boxes
    assert vol.weight.max() == 6 and (vol.weight == 0).mean() < 0.9
    # the table plane z = 0.05: the zero crossing lies between voxel layers 6 and 7 (voxel centres 0.04875 / 0.05625)
    col = vol.tsdf[2, 2]
    assert col[7] > 0 > col[5]
