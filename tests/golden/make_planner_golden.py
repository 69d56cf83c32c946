#!/usr/bin/env python
"""Golden fixtures for the planner post-processing from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_planner_golden.py

Imports vgn.detection_implicit with the shims of make_golden.py plus stub modules for packages its
transitive imports want but this path never executes (pyrender, urdfpy, open3d, skimage, pybullet,
matplotlib.pylab), then
  * calls the reference's own process() / bound() / select() on oracle.planner_oracle.seeded_volumes
    (stage fixtures: processed quality volume, grasp voxel indices, scores, quaternions, widths,
    translations -- in the reference's own output order), and
  * runs VGNImplicit.__call__ end to end (seeded network of make_golden.py, seeded TSDF) for one scene
    (e2e fixture: the grasps it returns with best=True, plus a strided sample of its raw volumes).
Inputs are regenerated from seeds by the tests; their checksums are stored.
"""
import os
import sys
from unittest.mock import MagicMock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

CASES = {   # tag -> (seed, plateau, weak, force_detection, max_filter_size)
    "c0": (0, False, False, False, 4),
    "c1": (1, True, False, False, 4),
    "c2": (2, False, True, False, 4),
    "c3": (2, False, True, True, 4),
    "c4": (3, True, True, True, 8),
    "c5": (4, False, False, True, 8),
}


def import_detection():
    import make_golden as M

    M.import_reference()
    for m in ["pyrender", "matplotlib.pylab", "urdfpy", "open3d", "skimage", "skimage.measure", "pybullet", "mcubes"]:
        s = MagicMock()
        s.__path__ = []
        sys.modules[m] = s
    import vgn.detection_implicit as D

    return D


def csum(a):
    a = np.ascontiguousarray(a).astype(np.float64)
    return np.array([a.sum(), np.abs(a).sum()])


def main():
    import torch

    from oracle import giga_oracle as O
    from oracle import planner_oracle as P

    D = import_detection()
    out = {}
    center = P.lattice_positions().view(40, 40, 40, 3)
    for tag, (seed, plateau, weak, fd, mfs) in CASES.items():
        tsdf, qual, rot, width = P.seeded_volumes(seed, plateau, weak)
        out[f"{tag}_cfg"] = np.array([seed, int(plateau), int(weak), int(fd), mfs])
        out[f"{tag}_in_checksum"] = np.stack([csum(tsdf), csum(qual), csum(rot), csum(width)])
        q, r, w = D.process(tsdf[None], qual.copy(), rot, width, out_th=0.5)
        q = D.bound(q, 0.3 / 40)
        grasps, scores = D.select(q.copy(), center, r, w, threshold=0.9, force_detection=fd, max_filter_size=mfs)
        nz = np.flatnonzero(q)
        out[f"{tag}_qvol_nz_index"] = nz.astype(np.int32)            # processed volume, sparse (most of it is masked to 0)
        out[f"{tag}_qvol_nz_value"] = q.reshape(-1)[nz]
        out[f"{tag}_scores"] = np.asarray(scores, np.float32)
        out[f"{tag}_trans"] = np.array([g.pose.translation for g in grasps], np.float32).reshape(-1, 3)
        out[f"{tag}_quat"] = np.array([g.pose.rotation.as_quat() for g in grasps], np.float64).reshape(-1, 4)
        out[f"{tag}_width"] = np.array([g.width for g in grasps], np.float32)
        print(tag, "grasps", len(grasps), "nonzero voxels", len(nz))

    # ---- end to end: the reference's VGNImplicit.__call__ on a seeded network + TSDF ----
    import vgn.networks as N

    torch.set_num_threads(8)
    net = N.get_network("giga")
    net.load_state_dict(P.planner_state_dict(O.seeded_state_dict(seed=1)))
    net.eval()
    for tag, seed, fd in (("e0", 5, False), ("e1", 6, True)):
        planner = object.__new__(D.VGNImplicit)      # __init__ would load a checkpoint from disk; same attributes set by hand
        planner.device = torch.device("cpu")
        planner.net = net
        planner.qual_th, planner.best, planner.force_detection, planner.out_th, planner.visualize = 0.9, True, fd, 0.5, False
        planner.resolution = 40
        planner.pos = P.lattice_positions()
        tsdf = P.seeded_volumes(seed)[0]
        state = type("State", (), {})()
        state.tsdf = tsdf[None]                      # np.ndarray branch of __call__ (detection_implicit.py:39-42)
        grasps, scores, _ = planner(state)
        qv, rv, wv = D.predict(tsdf[None], planner.pos, net, planner.device)
        out[f"{tag}_cfg"] = np.array([seed, int(fd)])
        out[f"{tag}_in_checksum"] = csum(tsdf)
        out[f"{tag}_scores"] = np.asarray(scores, np.float32)
        out[f"{tag}_trans"] = np.array([g.pose.translation for g in grasps], np.float32).reshape(-1, 3)
        out[f"{tag}_quat"] = np.array([g.pose.rotation.as_quat() for g in grasps], np.float64).reshape(-1, 4)
        out[f"{tag}_width"] = np.array([g.width for g in grasps], np.float32)
        out[f"{tag}_raw_qual"] = qv.reshape(-1)[::7].copy()
        out[f"{tag}_raw_width"] = wv.reshape(-1)[::7].copy()
        out[f"{tag}_raw_rot"] = rv.reshape(-1, 4)[::7].copy()
        print(tag, "grasps", len(grasps), "raw qual > 0.9:", int((qv > 0.9).sum()))
    path = os.path.join(HERE, "planner_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
