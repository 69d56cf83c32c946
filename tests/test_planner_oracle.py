"""CPU tests of the planner post-processing oracle (oracle/planner_oracle.py): pinned against scipy.ndimage itself
and against fixtures produced by the UNMODIFIED reference functions (tests/golden/make_planner_golden.py)."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import planner_oracle as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pgolden():
    return np.load(os.path.join(ROOT, "tests", "golden", "planner_golden.npz"), allow_pickle=False)


def _csum(a):
    a = np.ascontiguousarray(a).astype(np.float64)
    return np.array([a.sum(), np.abs(a).sum()])


def test_restatement_matches_scipy_bit_for_bit():
    ndimage = pytest.importorskip("scipy.ndimage")
    for seed, plateau in ((11, False), (12, True)):
        tsdf, qual, rot, width = P.seeded_volumes(seed, plateau)
        assert np.array_equal(ndimage.gaussian_filter(qual, sigma=1.0, mode="nearest"), P.gaussian_filter_nearest(qual))
        outside = tsdf > 0.5
        inside = np.logical_and(1e-3 < tsdf, tsdf < 0.5)
        assert np.array_equal(ndimage.binary_dilation(outside, iterations=2, mask=np.logical_not(inside)),
                              P.binary_dilation_masked(outside, 2, np.logical_not(inside)))
        sm = P.gaussian_filter_nearest(qual)
        for size in (3, 4, 8):
            assert np.array_equal(ndimage.maximum_filter(sm, size=size), P.maximum_filter_reflect(sm, size))


def test_oracle_matches_reference_golden(pgolden):
    """process()/bound()/select() of the unmodified reference on the same seeded volumes."""
    center = P.lattice_positions().view(40, 40, 40, 3).numpy()
    for tag in ("c0", "c1", "c2", "c3", "c4", "c5"):
        seed, plateau, weak, fd, mfs = (int(v) for v in pgolden[f"{tag}_cfg"])
        tsdf, qual, rot, width = P.seeded_volumes(seed, bool(plateau), bool(weak))
        assert np.allclose(np.stack([_csum(tsdf), _csum(qual), _csum(rot), _csum(width)]), pgolden[f"{tag}_in_checksum"], rtol=1e-12)
        idx, scores, r, w, qv = P.detect(tsdf[None], qual, rot, width, force_detection=bool(fd), max_filter_size=mfs)
        ref = np.zeros(64000, np.float32)
        ref[pgolden[f"{tag}_qvol_nz_index"]] = pgolden[f"{tag}_qvol_nz_value"]
        assert np.array_equal(qv.reshape(-1), ref), tag
        assert np.array_equal(scores, pgolden[f"{tag}_scores"]), tag          # same descending sequence
        # tie order is numpy's unstable quicksort in the reference: compare as sets of (score, position, width)
        ours = sorted((float(s), *map(float, center[tuple(i)]), float(x)) for s, i, x in zip(scores, idx, w))
        theirs = sorted((float(s), *map(float, t), float(x)) for s, t, x in zip(pgolden[f"{tag}_scores"], pgolden[f"{tag}_trans"], pgolden[f"{tag}_width"]))
        assert ours == theirs, tag
        if len(idx):
            pos2quat = {tuple(map(float, t)): q for t, q in zip(pgolden[f"{tag}_trans"], pgolden[f"{tag}_quat"])}
            for i, q in zip(idx, r):
                qn = q.astype(np.float64) / np.linalg.norm(q.astype(np.float64))
                assert np.allclose(qn, pos2quat[tuple(map(float, center[tuple(i)]))], atol=1e-12)


def test_bound_zero_limit_quirk():
    q = np.ones((40, 40, 40), np.float32)
    assert P.bound(q.copy(), 1.0).sum() == 0          # int(0.02 / 1.0) == 0 -> q[-0:] = 0 clears everything
    assert P.bound(q.copy(), 0.3 / 40).sum() == 36 * 36 * 33


def test_library_gaussian_weights_equal_numpy():
    """The fp64 kernel the device passes use (host helper of the C ABI) == scipy/numpy's, bit for bit, for the reference's sigma."""
    from giga_b200._lib import lib
    out = (C.c_double * 9)()
    assert lib.giga_gaussian_kernel1d(1.0, 4, out) == 0
    assert np.array_equal(np.array(out[:]), P.gaussian_kernel1d(1.0, 4))
    assert lib.giga_gaussian_kernel1d(1.0, 99, out) < 0
