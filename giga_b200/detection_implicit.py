"""Host-side mirror of the reference grasp planner (/root/reference/src/vgn/detection_implicit.py).

`VGNImplicit(model_path, model_type, ...)(state)` has the reference's constructor arguments, call
signature and return value `(grasps, scores, toc)`, so `scripts/sim_grasp_multiple.py` /
`vgn.experiments.clutter_removal.run` can use it unchanged.  Where the reference runs the network on the
GPU, copies 64,000 x 6 floats back and post-processes them with scipy.ndimage on the CPU
(predict :99-113, process :115-143, bound :87-97, select :146-174), this class makes ONE C-ABI call
(`giga_detect_host`): H2D of the 40^3 TSDF, encoder + grasp heads at the 40^3 query lattice, gaussian
smoothing / surface mask / width gate / border / NMS / sort on the device, D2H of the surviving grasps only.
There is no CPU path: the arithmetic is in libgiga_b200.so.
"""
from __future__ import annotations

import ctypes as C
import time

import numpy as np
import scipy.spatial.transform
import torch

from . import _lib
from ._lib import SelectParams, check, lib
from .networks import load_network

LOW_TH = 0.5   # detection_implicit.py:15


class Rotation(scipy.spatial.transform.Rotation):   # vgn/utils/transform.py:5-8
    @classmethod
    def identity(cls):
        return cls.from_quat([0.0, 0.0, 0.0, 1.0])


class Transform:
    """rotation (scipy Rotation) + translation (float64 ndarray), as vgn/utils/transform.py:11-47."""

    def __init__(self, rotation, translation):
        assert isinstance(rotation, scipy.spatial.transform.Rotation)
        self.rotation = rotation
        self.translation = np.asarray(translation, np.double)

    def as_matrix(self):
        m = np.eye(4)
        m[:3, :3] = self.rotation.as_matrix()
        m[:3, 3] = self.translation
        return m

    def to_list(self):
        return np.r_[self.rotation.as_quat(), self.translation]


class Grasp:   # vgn/grasp.py:9-18
    def __init__(self, pose, width):
        self.pose = pose
        self.width = width


def lattice(resolution: int = 40) -> torch.Tensor:
    """The query lattice of VGNImplicit.__init__ (detection_implicit.py:28-31): (1, res^3, 3) float32 on the CPU."""
    lin = torch.linspace(start=-0.5, end=0.5 - 1.0 / resolution, steps=resolution)
    x, y, z = torch.meshgrid(lin, lin, lin, indexing="ij")
    return torch.stack((x, y, z), dim=-1).float().reshape(1, resolution ** 3, 3)


def select_params(qual_th=0.9, out_th=0.5, force_detection=False, max_filter_size=4, voxel_size=0.3 / 40,
                  gaussian_filter_sigma=1.0, min_width=0.033, max_width=0.233, limit=(0.02, 0.02, 0.055)) -> SelectParams:
    """The arguments of process()/bound()/select() as a giga_select_params struct (bound's int(limit / voxel_size)
    is evaluated here, in Python float arithmetic like the reference)."""
    p = SelectParams()
    p.gaussian_sigma = float(gaussian_filter_sigma)
    p.min_width, p.max_width, p.out_th = min_width, max_width, out_th
    p.lim_x, p.lim_y, p.lim_z = (int(l / voxel_size) for l in limit)
    p.low_th, p.threshold = LOW_TH, qual_th
    p.force_detection, p.max_filter_size = int(bool(force_detection)), int(max_filter_size)
    return p


def detect_host(net, tsdf: np.ndarray, tsdf_process=None, params: SelectParams = None, K: int = 512):
    """giga_detect_host for B scenes: host TSDF(s) (B,40,40,40) float32 -> per scene the sorted grasps
    (count [B], score [B,K], index [B,K], rot [B,K,4], width [B,K]) as numpy arrays; re-runs with a larger K in the
    rare case a scene has more than K grasps."""
    eng = net._engine()
    if not getattr(eng, "lattice_set", False):
        pos = lattice().contiguous()
        check(lib.giga_ctx_set_lattice(eng.h, C.c_void_p(pos.data_ptr()), pos.shape[1]), "giga_ctx_set_lattice")
        eng.lattice_set = True
    params = params or select_params()
    tsdf = np.ascontiguousarray(tsdf, dtype=np.float32)
    if tsdf.ndim != 4 or tsdf.shape[1:] != (40, 40, 40):
        raise _lib.GigaError(f"tsdf must be (B,40,40,40), got {tsdf.shape}")
    B = tsdf.shape[0]
    tp = None
    if tsdf_process is not None:
        tp = np.ascontiguousarray(np.asarray(tsdf_process, dtype=np.float32).reshape(B, 40, 40, 40))
    ptr = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else C.c_void_p(0)
    while True:
        count = np.zeros(B, np.int32)
        score, index = np.zeros((B, K), np.float32), np.zeros((B, K), np.int32)
        rot, width = np.zeros((B, K, 4), np.float32), np.zeros((B, K), np.float32)
        stream = C.c_void_p(torch.cuda.current_stream(eng.device).cuda_stream)
        check(lib.giga_detect_host(eng.h, ptr(tsdf), ptr(tp), B, C.byref(params), K, ptr(count), ptr(score), ptr(index), ptr(rot),
                                   ptr(width), stream), "giga_detect_host")
        if int(count.max()) <= K:
            return count, score, index, rot, width
        K = int(count.max())


class VGNImplicit(object):
    """detection_implicit.py:17-85"""

    def __init__(self, model_path, model_type, best=False, force_detection=False, qual_th=0.9, out_th=0.5, visualize=False,
                 resolution=40, **kwargs):
        if not torch.cuda.is_available():
            raise _lib.GigaError("giga_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        if resolution != 40:
            raise _lib.GigaError("the planner kernels are built for the 40^3 grid the network takes (detection_implicit.py:100)")
        if visualize:
            raise NotImplementedError("mesh visualisation (vgn.utils.visual) is outside the hot path")
        self.device = torch.device("cuda")
        self.net = load_network(model_path, self.device, model_type=model_type) if model_path is not None else None
        self.qual_th = qual_th
        self.best = best
        self.force_detection = force_detection
        self.out_th = out_th
        self.visualize = visualize
        self.resolution = resolution
        self.pos = lattice(resolution)              # kept on the host: only the surviving voxels are looked up
        self._center = self.pos.view(resolution, resolution, resolution, 3).numpy()

    def __call__(self, state, scene_mesh=None, aff_kwargs={}):
        tsdf_process = state.tsdf_process if hasattr(state, "tsdf_process") else state.tsdf
        if isinstance(state.tsdf, np.ndarray):
            tsdf_vol = state.tsdf
            voxel_size = 0.3 / self.resolution
            size = 0.3
        else:
            tsdf_vol = state.tsdf.get_grid()
            voxel_size = tsdf_process.voxel_size
            tsdf_process = tsdf_process.get_grid()
            size = state.tsdf.size
        assert tsdf_vol.shape == (1, 40, 40, 40)    # predict(), detection_implicit.py:100

        tic = time.time()
        prm = select_params(qual_th=self.qual_th, out_th=self.out_th, force_detection=self.force_detection,
                            max_filter_size=8 if self.visualize else 4, voxel_size=voxel_size)
        same = tsdf_process is tsdf_vol or tsdf_process is state.tsdf
        count, score, index, rot, width = detect_host(self.net, tsdf_vol, None if same else tsdf_process, prm)
        n = int(count[0])
        grasps, scores = [], []
        for i in range(n):                           # select_index(), detection_implicit.py:177-185
            ijk = np.unravel_index(int(index[0, i]), (self.resolution,) * 3)
            ori = Rotation.from_quat(rot[0, i])
            grasps.append(Grasp(Transform(ori, self._center[ijk]), width[0, i]))
            scores.append(score[0, i])
        toc = time.time() - tic

        grasps, scores = np.asarray(grasps), np.asarray(scores)
        new_grasps = []
        if len(grasps) > 0:
            p = np.arange(len(grasps)) if self.best else np.random.permutation(len(grasps))
            for g in grasps[p]:
                pose = g.pose
                pose.translation = (pose.translation + 0.5) * size
                new_grasps.append(Grasp(pose, g.width * size))
            scores = scores[p]
        return new_grasps, scores, toc
