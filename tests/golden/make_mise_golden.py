"""Golden fixture for the MISE sweep, produced by the UNMODIFIED reference octree (mise.pyx compiled by oracle/build_ref.py) driven by the
loop of conv_onet/generation.py:127-143 on the deterministic field of tests/test_mise_oracle.py.
    python tests/golden/make_mise_golden.py   ->  tests/golden/mise_golden.npz"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import mise_oracle as M          # only its sweep() driver (= the reference loop) is used here
from oracle.build_ref import load_mise
from tests.test_mise_oracle import field

ref = load_mise()
assert ref is not None, "needs /root/reference + Cython"
out = {}
for key, (seed, r0, depth) in {"a": (11, 8, 2), "b": (12, 4, 3)}.items():
    grid, iters, total, _ = M.sweep(field(seed), r0, depth, 0.5, 0.1, mise_cls=ref.MISE)
    out[key + "_cfg"] = np.array([seed, r0, depth])
    out[key + "_grid"] = grid.astype(np.float32)       # the values are fp32 network outputs widened to double
    out[key + "_iters"] = np.array(iters)
    out[key + "_total"] = np.array(total)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "mise_golden.npz"), **out)
print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})

# ---- the whole reference pipeline: reference network (CPU) + reference Generator3D.generate_from_latent loop + reference MISE ----
from tests.golden.make_golden import import_reference   # noqa: E402

N = import_reference()
import torch                                                           # noqa: E402
import vgn.ConvONets.conv_onet.generation as G                         # noqa: E402
from oracle import giga_oracle as O                                    # noqa: E402

G.MISE = ref.MISE                                                      # the module's own `from ...libmise import MISE` needs the built extension
net = N.get_network("giga")
net.load_state_dict(O.seeded_state_dict(seed=1))
net.eval()
x, _, _ = O.seeded_inputs(1, 8, seed=21)
near = []


class Gen(G.Generator3D):
    def extract_mesh(self, occ_hat, c=None, stats_dict=dict()):       # stop before marching cubes: the sweep's result is the value grid
        return occ_hat

    def eval_points(self, p, c=None, **kw):
        v = super().eval_points(p, c, **kw)
        near.append(float((v - thr).abs().min()))
        return v


for key, (r0, steps, th, pad) in {"c": (8, 2, 0.5, 0.0), "d": (4, 3, 0.45, 0.1)}.items():
    thr = float(np.log(th) - np.log(1.0 - th))
    near.clear()
    gen = Gen(net, device=torch.device("cpu"), threshold=th, input_type="pointcloud", padding=pad, resolution0=r0, upsampling_steps=steps)
    with torch.no_grad():
        c = net.encode_inputs(x)
        grid = gen.generate_from_latent(c)
    out[key + "_cfg"] = np.array([r0, steps, th, pad], np.float64)
    out[key + "_grid"] = np.asarray(grid, np.float64).astype(np.float32)
    out[key + "_margin"] = np.array(min(near))                      # smallest |logit - threshold| over all evaluated points
    print(key, grid.shape, "min |v - thr| =", min(near))
out["net_x"] = x.numpy()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "mise_golden.npz"), **out)
