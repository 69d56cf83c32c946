"""TEST INFRASTRUCTURE ONLY -- CPU (numpy) restatement of the MISE octree that `Generator3D.generate_from_latent` drives.

Follows /root/reference/src/vgn/ConvONets/utils/libmise/mise.pyx (a Cython class; compiled from the reference sources into
oracle/_ref by oracle/build_ref.py and used to PIN this restatement: tests/test_mise_oracle.py compares, iteration by iteration,
the set of queried points and, at the end, to_dense() bit for bit):

    __cinit__          mise.pyx:42-85     resolution0^3 leaf voxels of size 2^depth, grid points at their corners
    query              mise.pyx:112-136   all grid points whose value is unknown
    update             mise.pyx:87-110    store values, then subdivide_voxels()
    subdivide_voxels   mise.pyx:184-232   a known point marks the LEAVES containing point + {-1,0}^3 (= every leaf whose closed box
                                          holds the point) next_to_positive (value >= threshold) / next_to_negative (value <= threshold);
                                          leaves below the finest level that are both are subdivided
    subdivide_voxel    mise.pyx:234-281   8 children, the 27 grid points at half spacing are added if new
    to_dense           mise.pyx:138-176   values on the finest lattice, gaps filled from the lower neighbour along x, then y, then z

The octree is restated densely on the finest lattice (what the CUDA kernels in giga_b200/csrc/mise.cuh do as well): `exists` / `known` /
`value` per grid point, `level` of the containing leaf per fine cell.  query() returns the points in lattice order, not in the
reference's insertion order: the result of the sweep does not depend on it (values are a function of position).
Only tests/ and __graft_entry__.smoke() may import this module.
"""
from __future__ import annotations

import itertools

import numpy as np


class MISE:
    def __init__(self, resolution_0: int, depth: int, threshold: float):
        self.resolution_0, self.depth, self.threshold = int(resolution_0), int(depth), float(threshold)
        self.voxel_size_0 = 1 << self.depth
        self.resolution = self.resolution_0 * self.voxel_size_0
        R, L = self.resolution, self.resolution + 1
        self.exists = np.zeros((L, L, L), bool)
        self.exists[:: self.voxel_size_0, :: self.voxel_size_0, :: self.voxel_size_0] = True
        self.known = np.zeros((L, L, L), bool)
        self.value = np.zeros((L, L, L), np.float64)
        self.level = np.zeros((R, R, R), np.int8)

    def query(self) -> np.ndarray:
        return np.argwhere(self.exists & ~self.known).astype(np.int64)

    def update(self, points: np.ndarray, values: np.ndarray) -> None:
        points = np.asarray(points, np.int64)
        assert points.shape[0] == values.shape[0] and points.shape[1] == 3
        if not self.exists[points[:, 0], points[:, 1], points[:, 2]].all():
            raise ValueError("Point not in grid!")
        self.value[points[:, 0], points[:, 1], points[:, 2]] = values
        self.known[points[:, 0], points[:, 1], points[:, 2]] = True
        self._subdivide_voxels()

    def _subdivide_voxels(self) -> None:
        R = self.resolution
        pos_pt = self.known & (self.value >= self.threshold)
        neg_pt = self.known & (self.value <= self.threshold)
        todo = []
        for lv in range(self.depth):                       # leaves of the finest level are never subdivided
            s = 1 << (self.depth - lv)
            leaf = self.level[::s, ::s, ::s] == lv          # origins of the level-lv leaves (a leaf's cells share its level)
            if not leaf.any():
                continue
            pos = np.zeros_like(leaf)
            neg = np.zeros_like(leaf)
            for i, j, k in itertools.product(range(s + 1), repeat=3):    # the (s+1)^3 lattice points of the closed box
                pos |= pos_pt[i : i + R : s, j : j + R : s, k : k + R : s]
                neg |= neg_pt[i : i + R : s, j : j + R : s, k : k + R : s]
            todo.append((lv, s, leaf & pos & neg))
        for lv, s, act in todo:                             # marking used the structure of the START of the call (mise.pyx:196-226)
            if not act.any():
                continue
            cells = act.repeat(s, 0).repeat(s, 1).repeat(s, 2)
            self.level[cells] = lv + 1
            ox, oy, oz = (a * s for a in np.nonzero(act))
            h = s >> 1
            for i, j, k in itertools.product(range(3), repeat=3):
                self.exists[ox + i * h, oy + j * h, oz + k * h] = True

    def to_dense(self) -> np.ndarray:
        out = np.where(self.exists, self.value, np.nan)
        n = self.resolution + 1
        for i in range(1, n):
            m = np.isnan(out[i])
            out[i][m] = out[i - 1][m]
        for j in range(1, n):
            m = np.isnan(out[:, j])
            out[:, j][m] = out[:, j - 1][m]
        for k in range(1, n):
            m = np.isnan(out[:, :, k])
            out[:, :, k][m] = out[:, :, k - 1][m]
        assert not np.isnan(out).any()
        return out


def sweep(eval_points, resolution0: int = 16, upsampling_steps: int = 3, threshold: float = 0.5, padding: float = 0.1, mise_cls=MISE):
    """generate_from_latent's MISE loop (conv_onet/generation.py:102-143) with `eval_points(pointsf float32 (n,3)) -> logits float32 (n,)`
    standing in for Generator3D.eval_points.  Returns (value_grid float64, iterations, points evaluated, list of per-iteration point sets)."""
    thr = np.log(threshold) - np.log(1.0 - threshold)
    box_size = 1 + padding
    ex = mise_cls(resolution0, upsampling_steps, thr)
    points = ex.query()
    iters, total, sets = 0, 0, []
    while points.shape[0] != 0:
        pointsf = points / ex.resolution
        pointsf = (box_size * (pointsf - 0.5)).astype(np.float32)          # torch.FloatTensor(pointsf)
        values = np.asarray(eval_points(pointsf), np.float32).astype(np.float64)
        ex.update(points, values)
        sets.append({tuple(p) for p in points.tolist()})
        iters += 1
        total += points.shape[0]
        points = ex.query()
    return ex.to_dense(), iters, total, sets
