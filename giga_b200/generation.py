"""Host-side mirror of the reference mesh generator's occupancy sweep
(/root/reference/src/vgn/ConvONets/conv_onet/generation.py: `Generator3D`).

`Generator3D(model, ...)` takes the reference's constructor arguments; `generate_mesh(data)` / `generate_from_latent(c)` follow
generation.py:65-143.  What the reference does with a Python loop that ships every MISE query batch to the GPU and every value back to a
Cython octree on the host (`MISE.query` -> `eval_points` -> `MISE.update`, generation.py:127-143, utils/libmise/mise.pyx) is ONE C-ABI call
here (`giga_mise_sweep`): octree bookkeeping, point arithmetic and TSDF-head evaluation stay on the device; the dense value grid
(`mesh_extractor.to_dense()`) comes back.  `upsampling_steps == 0` evaluates the regular grid of generation.py:117-125 directly.

Mesh extraction itself (marching cubes = utils/libmcubes, trimesh, optional simplification / refinement; generation.py:360-428) is CPU
geometry post-processing outside the hot path: `extract_mesh` uses the reference's own `libmcubes` + `trimesh` when they are importable
(a GIGA environment has them) and raises otherwise; `generate_value_grid` is always available.  There is no CPU path for the sweep.
"""
from __future__ import annotations

import ctypes as C
import time

import numpy as np
import torch

from . import _lib
from ._lib import check, lib


class Generator3D(object):
    """generation.py:20-64 (same arguments, same defaults)."""

    def __init__(self, model, points_batch_size=100000, threshold=0.5, refinement_step=0, device=None, resolution0=16, upsampling_steps=3,
                 with_normals=False, padding=0.1, sample=False, input_type=None, vol_info=None, vol_bound=None, simplify_nfaces=None):
        self.model = model.to(device)
        self.points_batch_size = points_batch_size
        self.refinement_step = refinement_step
        self.threshold = threshold
        self.device = device
        self.resolution0 = resolution0
        self.upsampling_steps = upsampling_steps
        self.with_normals = with_normals
        self.input_type = input_type
        self.padding = padding
        self.sample = sample
        self.simplify_nfaces = simplify_nfaces
        self.vol_bound = vol_bound
        if vol_bound is not None or vol_info is not None:
            raise NotImplementedError("sliding-window crops (pointcloud_crop) are not used by GIGA and not built here")

    # -- the sweep ---------------------------------------------------------------------------------------------------------
    def generate_value_grid(self, c, stats_dict=None) -> torch.Tensor:
        """The occupancy-logit grid `generate_from_latent` hands to `extract_mesh`: (n, n, n) float32 on the device,
        n = resolution0 * 2**upsampling_steps + 1 (MISE) or resolution0 (upsampling_steps == 0)."""
        stats_dict = {} if stats_dict is None else stats_dict
        net = self.model
        eng = net._engine()
        planes = net._packed(c)
        if planes.shape[1] != 1:
            raise _lib.GigaError("Generator3D processes one scene at a time (generation.py:90-95)")
        box_size = 1 + self.padding
        t0 = time.time()
        if self.upsampling_steps == 0:                       # generation.py:117-125: regular grid, make_3d_grid (common.py:148-167)
            nx = self.resolution0
            lin = torch.linspace(-0.5, 0.5, nx)
            gx, gy, gz = torch.meshgrid(lin, lin, lin, indexing="ij")
            pointsf = (box_size * torch.stack([gx, gy, gz], dim=-1).reshape(1, -1, 3)).to(eng.device)
            grid = net.decode_occ(pointsf, c).logits.reshape(nx, nx, nx)
            stats_dict["points evaluated"] = nx ** 3
        else:
            n = (self.resolution0 << self.upsampling_steps) + 1
            grid = torch.empty((n, n, n), device=eng.device, dtype=torch.float32)
            threshold = np.log(self.threshold) - np.log(1. - self.threshold)      # generation.py:110
            stats = (C.c_int * 2)()
            check(lib.giga_mise_sweep(eng.h, C.c_void_p(planes.data_ptr()), int(self.resolution0), int(self.upsampling_steps), float(threshold),
                                      float(box_size), C.c_void_p(grid.data_ptr()), stats,
                                      C.c_void_p(torch.cuda.current_stream(eng.device).cuda_stream)), "giga_mise_sweep")
            stats_dict["mise iterations"], stats_dict["points evaluated"] = int(stats[0]), int(stats[1])
        stats_dict["time (eval points)"] = time.time() - t0
        return grid

    def generate_from_latent(self, c=None, stats_dict={}, **kwargs):
        """generation.py:102-148"""
        value_grid = self.generate_value_grid(c, stats_dict).double().cpu().numpy()      # the reference's grid is float64 on the host
        return self.extract_mesh(value_grid, c, stats_dict=stats_dict)

    def generate_mesh(self, data, return_stats=True):
        """generation.py:65-100"""
        self.model.eval()
        stats_dict = {}
        inputs = data.get("inputs", torch.empty(1, 0)).to(self.device)
        t0 = time.time()
        with torch.no_grad():
            c = self.model.encode_inputs(inputs)
        stats_dict["time (encode inputs)"] = time.time() - t0
        mesh = self.generate_from_latent(c, stats_dict=stats_dict)
        return (mesh, stats_dict) if return_stats else mesh

    def eval_points(self, p, c=None, **kwargs):
        """generation.py:326-358 (the non-crop branch): occupancy logits of the TSDF head at p (n, 3), on the CPU like the reference."""
        with torch.no_grad():
            return self.model.decode_occ(p.unsqueeze(0).to(self.device), c).logits.squeeze(0).cpu()

    # -- CPU geometry post-processing (the reference's own utilities, when present) ---------------------------------------------
    def extract_mesh(self, occ_hat, c=None, stats_dict=dict()):
        """generation.py:360-428 without normals / refinement / simplification: marching cubes on the padded grid, vertices mapped back to
        the unit cube.  Needs the reference's `libmcubes` extension and `trimesh`."""
        try:
            import trimesh
            from vgn.ConvONets.utils import libmcubes
        except Exception as e:   # noqa: BLE001
            raise NotImplementedError("mesh extraction (marching cubes + trimesh) is CPU post-processing outside the hot path; install the "
                                      "reference's vgn.ConvONets.utils.libmcubes and trimesh, or use generate_value_grid()") from e
        n_x, n_y, n_z = occ_hat.shape
        box_size = 1 + self.padding
        threshold = np.log(self.threshold) - np.log(1. - self.threshold)
        t0 = time.time()
        occ_hat_padded = np.pad(occ_hat, 1, "constant", constant_values=-1e6)
        vertices, triangles = libmcubes.marching_cubes(occ_hat_padded, threshold)
        stats_dict["time (marching cubes)"] = time.time() - t0
        vertices -= 0.5
        vertices -= 1
        vertices /= np.array([n_x - 1, n_y - 1, n_z - 1])
        vertices = box_size * (vertices - 0.5)
        return trimesh.Trimesh(vertices, triangles, vertex_normals=None, process=False)
