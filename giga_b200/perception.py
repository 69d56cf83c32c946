"""Host mirror of `vgn/perception.py` for the hot path's upstream neighbour: `CameraIntrinsic` (:10-59),
`TSDFVolume` (:65-118) and `create_tsdf` (:121-126) with the reference's names, arguments and return values.  The arithmetic the
reference delegates to Open3D (`UniformTSDFVolume.integrate`, `extract_voxel_grid`) and the per-voxel Python loop of `get_grid` run as
CUDA kernels (csrc/tsdf.cuh) on volumes that stay on the device: `get_grid_device()` hands the (1, R, R, R) network input straight to
`net(...)` / `VGNImplicit` without the reference's device -> host -> device round trip; `get_grid()` returns the numpy array the
reference returns.  No CPU path: without the CUDA library / a GPU every call raises."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, lib


class CameraIntrinsic(object):
    """vgn/perception.py:10-59"""

    def __init__(self, width, height, fx, fy, cx, cy):
        self.width = width
        self.height = height
        self.K = np.array([[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]])

    fx = property(lambda self: self.K[0, 0])
    fy = property(lambda self: self.K[1, 1])
    cx = property(lambda self: self.K[0, 2])
    cy = property(lambda self: self.K[1, 2])

    def to_dict(self):
        return {"width": self.width, "height": self.height, "K": self.K.flatten().tolist()}

    @classmethod
    def from_dict(cls, data):
        return cls(width=data["width"], height=data["height"], fx=data["K"][0], fy=data["K"][4], cx=data["K"][2], cy=data["K"][5])


def _matrix(extrinsic) -> np.ndarray:
    m = extrinsic.as_matrix() if hasattr(extrinsic, "as_matrix") else extrinsic      # vgn.utils.transform.Transform or a 4x4
    m = np.ascontiguousarray(np.asarray(m, np.float64))
    if m.shape != (4, 4):
        raise _lib.GigaError(f"extrinsic must be a 4x4 transform (T_eye_task), got {m.shape}")
    return m


class TSDFVolume(object):
    """Integration of multiple depth images using a TSDF (vgn/perception.py:65-118)."""

    def __init__(self, size, resolution, device=None):
        from .training import _engine
        self.size = size
        self.resolution = resolution
        self.voxel_size = self.size / self.resolution
        self.sdf_trunc = 4 * self.voxel_size
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
        self._eng = _engine(dev)
        self._device = self._eng.device
        R = resolution
        self._tsdf = torch.zeros((R, R, R), device=self._device, dtype=torch.float32)
        self._weight = torch.zeros((R, R, R), device=self._device, dtype=torch.float32)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self._device).cuda_stream)

    def integrate(self, depth_img, intrinsic, extrinsic):
        """depth_img: (H, W) depth in metres (numpy or a tensor already on the device); intrinsic: CameraIntrinsic; extrinsic: T_eye_task"""
        self.integrate_views(depth_img[None], intrinsic, [extrinsic])

    def integrate_views(self, depth_imgs, intrinsic, extrinsics):
        """all views of a scan in one call (create_tsdf's loop, perception.py:123-125): depth_imgs (n, H, W)"""
        d = torch.as_tensor(np.asarray(depth_imgs) if not torch.is_tensor(depth_imgs) else depth_imgs)
        d = d.to(device=self._device, dtype=torch.float32).contiguous()
        if d.dim() != 3 or tuple(d.shape[1:]) != (int(intrinsic.height), int(intrinsic.width)):
            raise _lib.GigaError(f"depth images must be (n, {intrinsic.height}, {intrinsic.width}), got {tuple(d.shape)}")
        n = d.shape[0]
        E = np.ascontiguousarray(np.stack([_matrix(e) for e in extrinsics]))
        if E.shape[0] != n:
            raise _lib.GigaError("one extrinsic per depth image")
        check(lib.giga_tsdf_integrate(self._eng.h, C.c_void_p(self._tsdf.data_ptr()), C.c_void_p(self._weight.data_ptr()), self.resolution,
                                      float(self.size), float(self.sdf_trunc), C.c_void_p(d.data_ptr()), n, int(intrinsic.width), int(intrinsic.height),
                                      float(intrinsic.fx), float(intrinsic.fy), float(intrinsic.cx), float(intrinsic.cy),
                                      E.ctypes.data_as(C.POINTER(C.c_double)), 1.0, 2.0, self._stream()), "giga_tsdf_integrate")

    def get_grid_device(self) -> torch.Tensor:
        """(1, R, R, R) float32 on the device: the network input"""
        R = self.resolution
        grid = torch.empty((1, R, R, R), device=self._device, dtype=torch.float32)
        check(lib.giga_tsdf_grid(self._eng.h, C.c_void_p(self._tsdf.data_ptr()), C.c_void_p(self._weight.data_ptr()), R,
                                 C.c_void_p(grid.data_ptr()), self._stream()), "giga_tsdf_grid")
        return grid

    def get_grid(self):
        """perception.py:106-115: (1, R, R, R) float32 numpy array"""
        return self.get_grid_device().cpu().numpy()

    def get_cloud(self):
        raise NotImplementedError("get_cloud (Open3D point-cloud extraction, perception.py:117-118) is outside the hot path (SURVEY.md section 2)")


def create_tsdf(size, resolution, depth_imgs, intrinsic, extrinsics):
    """perception.py:121-126; `extrinsics[i]` = a 7-vector [qx qy qz qw x y z] (Transform.from_list), a Transform or a 4x4"""
    tsdf = TSDFVolume(size, resolution)
    mats = []
    for e in extrinsics:
        e_arr = np.asarray(e, np.float64) if not hasattr(e, "as_matrix") else None
        if e_arr is not None and e_arr.shape == (7,):
            from scipy.spatial.transform import Rotation
            m = np.eye(4)
            m[:3, :3] = Rotation.from_quat(e_arr[:4]).as_matrix()
            m[:3, 3] = e_arr[4:]
            mats.append(m)
        else:
            mats.append(_matrix(e))
    tsdf.integrate_views(depth_imgs, intrinsic, mats)
    return tsdf
