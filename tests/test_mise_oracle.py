"""CPU: the numpy restatement of MISE (oracle/mise_oracle.py) against the UNMODIFIED reference octree
(/root/reference/src/vgn/ConvONets/utils/libmise/mise.pyx compiled into oracle/_ref by oracle/build_ref.py) and against the committed
golden fixture (tests/golden/mise_golden.npz, made with the compiled reference by tests/golden/make_mise_golden.py)."""
import os

import numpy as np
import pytest

from oracle import mise_oracle as M
from oracle.build_ref import load_mise

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def field(seed):
    """A deterministic occupancy-logit field with a curved iso-surface, thin sheets and exact threshold hits."""
    rng = np.random.RandomState(seed)
    c = rng.uniform(-0.2, 0.2, size=(3, 3)).astype(np.float32)
    r = rng.uniform(0.12, 0.3, size=3).astype(np.float32)

    def f(p):
        p = np.asarray(p, np.float32)
        d = np.stack([np.linalg.norm(p - c[i], axis=1) - r[i] for i in range(3)]).min(0)
        v = (-8.0 * d + 0.5 * np.sin(23.0 * p[:, 0]) * np.cos(17.0 * p[:, 1])).astype(np.float32)
        v[np.abs(v) < 0.05] = 0.0          # values exactly on the threshold (>= and <= both hold)
        return v
    return f


def sweep_pair(seed, r0, depth, th=0.5, padding=0.1):
    ref_mod = load_mise()
    if ref_mod is None:
        pytest.skip("oracle/_ref/mise not built (no /root/reference or Cython here)")
    f = field(seed)
    a = M.sweep(f, r0, depth, th, padding, mise_cls=M.MISE)
    b = M.sweep(f, r0, depth, th, padding, mise_cls=ref_mod.MISE)
    return a, b


@pytest.mark.parametrize("seed,r0,depth", [(0, 4, 2), (1, 8, 2), (2, 16, 3), (3, 3, 3), (4, 8, 1)])
def test_oracle_equals_compiled_reference(seed, r0, depth):
    (ga, ia, na, sa), (gb, ib, nb, sb) = sweep_pair(seed, r0, depth)
    assert ia == ib and na == nb
    for x, y in zip(sa, sb):
        assert x == y                       # the same grid points are queried in every iteration
    assert ga.shape == gb.shape == ((r0 << depth) + 1,) * 3
    assert np.array_equal(ga, gb)           # bit-exact dense grid


def test_threshold_on_other_levels():
    th = 0.3
    (ga, ia, na, _), (gb, ib, nb, _) = sweep_pair(7, 8, 2, th=th, padding=0.0)
    assert (ia, na) == (ib, nb) and np.array_equal(ga, gb)


def test_golden_fixture():
    g = np.load(os.path.join(ROOT, "tests", "golden", "mise_golden.npz"))
    for key in ("a", "b"):
        seed, r0, depth = (int(v) for v in g[key + "_cfg"])
        grid, iters, total, _ = M.sweep(field(seed), r0, depth, 0.5, 0.1)
        assert iters == int(g[key + "_iters"]) and total == int(g[key + "_total"])
        assert np.array_equal(grid.astype(np.float32), g[key + "_grid"])
