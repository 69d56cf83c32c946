"""Host-side mirror of the reference VGN grasp detector (/root/reference/src/vgn/detection.py).

`VGN(model_path, model_type, ...)(state)` has the reference's constructor arguments, call signature and return value
`(grasps, scores, toc)`.  predict (:95-109: H2D -> ConvNet -> D2H of three 40^3 volumes) + process (:111-145) + bound (:83-93) + select
(:147-175, scipy.ndimage on the host) are two C-ABI calls here: giga_vgn_forward (csrc/vgn.cuh) and giga_select_grasps (the same planner
kernels the implicit detector uses: the two `process`/`select` implementations differ only in the width gate, 1.33..9.33 voxels here).
"""
from __future__ import annotations

import time

import numpy as np
import torch

from . import _lib
from .detection_implicit import Grasp, Rotation, Transform, select_params
from .networks import load_network


class VGN(object):
    """detection.py:26-81"""

    def __init__(self, model_path, model_type, best=False, force_detection=False, qual_th=0.9, out_th=0.5, visualize=False):
        if not torch.cuda.is_available():
            raise _lib.GigaError("giga_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        if visualize:
            raise NotImplementedError("mesh visualisation (vgn.utils.visual) is outside the hot path")
        self.device = torch.device("cuda")
        self.net = load_network(model_path, self.device, model_type=model_type) if model_path is not None else None
        if self.net is not None:
            self.net.eval()
        self.qual_th = qual_th
        self.best = best
        self.force_detection = force_detection
        self.out_th = out_th
        self.visualize = visualize

    def detect(self, tsdf_vol: np.ndarray, voxel_size: float, K: int = 512):
        """predict + process + bound + select on the device -> (count, score, index, rot, width) numpy arrays for the scene."""
        assert tsdf_vol.shape == (1, 40, 40, 40)     # predict(), detection.py:96
        dev = next(self.net.parameters()).device
        t = torch.from_numpy(np.ascontiguousarray(tsdf_vol, dtype=np.float32)).to(dev)
        prm = select_params(qual_th=self.qual_th, out_th=self.out_th, force_detection=self.force_detection, max_filter_size=4, voxel_size=voxel_size,
                            min_width=1.33, max_width=9.33)      # process() defaults, detection.py:117-118 (widths in voxels)
        with torch.no_grad():
            qual, rot, width = self.net.forward_flat(t)
            while True:
                out = _select(self.net, t, qual, rot, width, prm, K)
                if int(out[0][0]) <= K:
                    return [o.cpu().numpy() for o in out]
                K = int(out[0][0])

    def __call__(self, state, scene_mesh=None, aff_kwargs={}):
        if isinstance(state.tsdf, np.ndarray):
            tsdf_vol = state.tsdf
            voxel_size = 0.3 / 40
        else:
            tsdf_vol = state.tsdf.get_grid()
            voxel_size = state.tsdf.voxel_size
        tic = time.time()
        count, score, index, rot, width = self.detect(tsdf_vol, voxel_size)
        n = int(count[0])
        grasps, scores = [], []
        for i in range(n):                           # select_index(), detection.py:177-183
            ijk = np.unravel_index(int(index[0, i]), (40, 40, 40))
            grasps.append(Grasp(Transform(Rotation.from_quat(rot[0, i]), np.array(ijk, dtype=np.float64)), width[0, i]))
            scores.append(score[0, i])
        toc = time.time() - tic
        grasps, scores = np.asarray(grasps), np.asarray(scores)
        if len(grasps) > 0:
            p = np.arange(len(grasps)) if self.best else np.random.permutation(len(grasps))
            out = []
            for g in grasps[p]:                      # from_voxel_coordinates(), grasp.py:27-31
                pose = g.pose
                pose.translation = pose.translation * voxel_size
                out.append(Grasp(pose, g.width * voxel_size))
            grasps, scores = out, scores[p]
        return grasps, scores, toc


def _select(net, tsdf, qual, rot, width, prm, K):
    import ctypes as C

    from ._lib import check, lib
    eng = net._engine()
    B = tsdf.shape[0]
    mk = lambda shape, dt: torch.zeros(shape, device=eng.device, dtype=dt)
    count, score, index = mk((B,), torch.int32), mk((B, K), torch.float32), mk((B, K), torch.int32)
    orot, owidth = mk((B, K, 4), torch.float32), mk((B, K), torch.float32)
    ptr = lambda a: C.c_void_p(a.data_ptr())
    check(lib.giga_select_grasps(eng.h, ptr(tsdf), ptr(qual), ptr(rot), ptr(width), B, C.byref(prm), K, ptr(count), ptr(score), ptr(index),
                                 ptr(orot), ptr(owidth), C.c_void_p(0), C.c_void_p(torch.cuda.current_stream(eng.device).cuda_stream)),
          "giga_select_grasps")
    return count, score, index, orot, owidth
