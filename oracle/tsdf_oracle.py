"""CPU oracle for the TSDF integration / read-out (TEST INFRASTRUCTURE, not product).  SURVEY.md 8f rank 2.

PARITY UNPINNED.  The reference side of this row is three lines of glue around a third-party library that is absent from /root/reference
and from this image (no network): `vgn/perception.py:65-115` `TSDFVolume` = `open3d.pipelines.integration.UniformTSDFVolume`
(open3d==0.12.0, requirements.txt:5) -- `integrate` (:79-104: RGBDImage.create_from_color_and_depth(depth_scale=1, depth_trunc=2,
convert_rgb_to_intensity=False) + volume.integrate) and `get_grid` (:106-115: extract_voxel_grid, colour channel 0 scattered into a
(1, R, R, R) float32 grid).  The reference ships no tests, golden vectors or fixtures for it.  What follows is a
numpy float32 restatement of Open3D 0.12's published algorithm:
  * cpp/open3d/geometry/ImageFactory.cpp  Image::ConvertDepthToFloatImage: d /= depth_scale; d >= depth_trunc -> 0;
    Image::CreateDepthToCameraDistanceMultiplierFloatImage: sqrt(((j - cx) / fx)^2 + ((i - cy) / fy)^2 + 1)
  * cpp/open3d/pipelines/integration/UniformTSDFVolume.cpp  IntegrateWithDepthToCameraDistanceMultiplier: per (x, y) column the voxel
    centre (half + length * x, half + length * y, half) is transformed once by the float extrinsic, then marched along z by adding the
    scaled third column; projection u = X fx / Z + cx + 0.5 (truncated to int), bounds [0.0001, dim - 0.0001); sdf = (d - Z) * multiplier;
    sdf > -trunc -> tsdf = min(1, sdf / trunc) folded into the running average with weight + 1
  * same file, ExtractVoxelGrid: voxels with weight != 0 and -0.98 <= tsdf < 0.98 get colour (tsdf + 1) / 2 (double)
The floating-point ORDER of the 4x4 transform (Eigen's vectorised product) is an assumption: row sums left to right, no fused
multiply-add.  Tests pin the CUDA kernel to this restatement bit for bit and check analytic properties of the restatement itself
(a fronto-parallel wall gives the closed-form TSDF ramp); agreement with Open3D's binaries cannot be shown here.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def depth_to_float(depth: np.ndarray, depth_scale: float = 1.0, depth_trunc: float = 2.0) -> np.ndarray:
    d = depth.astype(np.float32) / f32(depth_scale)
    d[d >= f32(depth_trunc)] = 0.0
    return d


def distance_multiplier(width: int, height: int, fx, fy, cx, cy) -> np.ndarray:
    xx = (np.arange(width, dtype=np.float32) - f32(cx)) * (f32(1.0) / f32(fx))
    yy = (np.arange(height, dtype=np.float32) - f32(cy)) * (f32(1.0) / f32(fy))
    return np.sqrt(xx[None, :] * xx[None, :] + yy[:, None] * yy[:, None] + f32(1.0)).astype(np.float32)


class TSDFVolume:
    """vgn/perception.py:65-115 over the restated UniformTSDFVolume (origin 0, NoColor)."""

    def __init__(self, size: float, resolution: int):
        self.size, self.resolution = size, resolution
        self.voxel_size = self.size / self.resolution
        self.sdf_trunc = 4 * self.voxel_size
        R = resolution
        self.tsdf = np.zeros((R, R, R), np.float32)       # [x][y][z]
        self.weight = np.zeros((R, R, R), np.float32)

    def integrate(self, depth_img, intrinsic, extrinsic):
        """intrinsic: object with width, height, fx, fy, cx, cy; extrinsic: 4x4 (or an object with as_matrix()) = T_eye_task"""
        E = np.asarray(extrinsic.as_matrix() if hasattr(extrinsic, "as_matrix") else extrinsic, np.float64).astype(np.float32)
        W, H = int(intrinsic.width), int(intrinsic.height)
        fx, fy, cx, cy = (f32(v) for v in (intrinsic.fx, intrinsic.fy, intrinsic.cx, intrinsic.cy))
        d = depth_to_float(np.asarray(depth_img), 1.0, 2.0)
        mult = distance_multiplier(W, H, fx, fy, cx, cy)
        R = self.resolution
        vl = f32(self.size / R)
        half = vl * f32(0.5)
        trunc = f32(self.sdf_trunc)
        trunc_inv = f32(1.0) / trunc
        Es = E * vl
        safe_w, safe_h = f32(W) - f32(0.0001), f32(H) - f32(0.0001)
        px = (half + vl * np.arange(R, dtype=np.float32))[:, None] * np.ones((1, R), np.float32)
        py = (half + vl * np.arange(R, dtype=np.float32))[None, :] * np.ones((R, 1), np.float32)
        pz = half
        cam = [((E[i, 0] * px + E[i, 1] * py) + E[i, 2] * pz) + E[i, 3] for i in range(3)]      # float32, left to right
        for z in range(R):
            X, Y, Z = cam
            with np.errstate(divide="ignore", invalid="ignore"):
                u_f = X * fx / Z + cx + f32(0.5)
                v_f = Y * fy / Z + cy + f32(0.5)
            ok = (Z > 0) & (u_f >= f32(0.0001)) & (u_f < safe_w) & (v_f >= f32(0.0001)) & (v_f < safe_h)
            u = np.where(ok, u_f, 0).astype(np.int32)
            v = np.where(ok, v_f, 0).astype(np.int32)
            dd = d[v, u]
            ok &= dd > 0
            sdf = (dd - Z) * mult[v, u]
            ok &= sdf > -trunc
            t = np.minimum(f32(1.0), sdf * trunc_inv)
            w = self.weight[:, :, z]
            new = (self.tsdf[:, :, z] * w + t) / (w + f32(1.0))
            self.tsdf[:, :, z] = np.where(ok, new, self.tsdf[:, :, z])
            self.weight[:, :, z] = np.where(ok, w + f32(1.0), w)
            cam = [cam[i] + Es[i, 2] for i in range(3)]

    def get_grid(self) -> np.ndarray:
        f, w = self.tsdf, self.weight
        keep = (w != 0) & (f < f32(0.98)) & (f >= f32(-0.98))
        c = ((f.astype(np.float64) + 1.0) * 0.5).astype(np.float32)
        return np.where(keep, c, f32(0.0))[None].astype(np.float32)


def create_tsdf(size, resolution, depth_imgs, intrinsic, extrinsics):
    """vgn/perception.py:121-126 with 4x4 extrinsics"""
    tsdf = TSDFVolume(size, resolution)
    for i in range(depth_imgs.shape[0]):
        tsdf.integrate(depth_imgs[i], intrinsic, extrinsics[i])
    return tsdf


# ------------------------------------------------------------------------------------------------------
# seeded synthetic scenes: depth images of a table plane with boxes, rendered analytically (pinhole, z-depth)
# ------------------------------------------------------------------------------------------------------
class Intrinsic:
    def __init__(self, width, height, fx, fy, cx, cy):
        self.width, self.height, self.fx, self.fy, self.cx, self.cy = width, height, fx, fy, cx, cy


def look_at(eye, target, up=(0.0, 0.0, 1.0)) -> np.ndarray:
    """T_eye_task as a 4x4 (camera looks along +z, x right, y down), the convention of vgn.perception.camera_on_sphere"""
    eye, target, up = (np.asarray(a, np.float64) for a in (eye, target, up))
    fwd = target - eye
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    Rm = np.stack([right, down, fwd])            # rows: camera axes in task coordinates
    T = np.eye(4)
    T[:3, :3] = Rm
    T[:3, 3] = -Rm @ eye
    return T


def render_depth(intr: Intrinsic, T: np.ndarray, boxes, table_z=0.05, noise=0.0, rng=None) -> np.ndarray:
    """z-depth image of the plane z = table_z and axis-aligned boxes [(lo, hi), ...] in task coordinates"""
    H, W = intr.height, intr.width
    j, i = np.meshgrid(np.arange(W), np.arange(H))
    dirs_c = np.stack([(j - intr.cx) / intr.fx, (i - intr.cy) / intr.fy, np.ones_like(j, np.float64)], -1)   # z-depth parametrisation
    Rm, t = T[:3, :3], T[:3, 3]
    eye = -Rm.T @ t
    dirs = dirs_c @ Rm                            # task-frame ray directions (per unit camera z)
    depth = np.full((H, W), np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        s = (table_z - eye[2]) / dirs[..., 2]
        depth = np.where(s > 0, s, depth)
        for lo, hi in boxes:
            lo, hi = np.asarray(lo, np.float64), np.asarray(hi, np.float64)
            t0 = (lo - eye) / dirs
            t1 = (hi - eye) / dirs
            tn = np.nanmax(np.minimum(t0, t1), -1)
            tf = np.nanmin(np.maximum(t0, t1), -1)
            hit = (tn <= tf) & (tn > 0)
            depth = np.where(hit & (tn < depth), tn, depth)
    depth = np.where(np.isfinite(depth), depth, 0.0)
    if noise > 0:
        depth = depth + (rng or np.random.default_rng(0)).normal(0, noise, depth.shape) * (depth > 0)
    return depth.astype(np.float32)


def seeded_scene(seed: int, n_views: int = 6, size: float = 0.3, width: int = 640, height: int = 480):
    """-> (depth_imgs [n][H][W] float32, Intrinsic, extrinsics [n][4][4] float64): the camera set-up of simulation.py:145-186
    (intrinsic 540/540/320/240, cameras on a circle of radius 2 * size at polar angle pi / 6 around the workspace centre)"""
    rng = np.random.default_rng(seed)
    intr = Intrinsic(width, height, 540.0 * width / 640, 540.0 * height / 480, 320.0 * width / 640, 240.0 * height / 480)
    boxes = []
    for _ in range(int(rng.integers(1, 5))):
        c = rng.uniform(0.06, size - 0.06, 2)
        h = rng.uniform(0.03, 0.15)
        w = rng.uniform(0.015, 0.04, 2)
        boxes.append(((c[0] - w[0], c[1] - w[1], 0.05), (c[0] + w[0], c[1] + w[1], 0.05 + h)))
    origin = np.array([size / 2, size / 2, 0.0])
    r, theta = 2.0 * size, np.pi / 6.0
    Ts, imgs = [], []
    for k in range(n_views):
        phi = 2.0 * np.pi * k / n_views
        eye = origin + r * np.array([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)])
        T = look_at(eye, origin)
        Ts.append(T)
        imgs.append(render_depth(intr, T, boxes, noise=0.0005 * (seed % 2), rng=rng))
    return np.stack(imgs), intr, np.stack(Ts)
