#!/usr/bin/env python
"""Generate golden fixtures by running the UNMODIFIED reference (GIGA @ /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference is imported in-process with three shims:
  * a pure-torch stand-in for the un-vendored third-party `torch_scatter`
    (torch-scatter 2.0.6, environment.yaml:145) -- scatter_mean restated from its
    published semantics,
  * `numpy.int = int` (ConvONets/utils/binvox_rw.py:206 uses the removed alias),
  * stub modules for matplotlib / mpl_toolkits / trimesh (imported transitively,
    never executed on this path).
Parameters and inputs come from oracle.giga_oracle.seeded_state_dict/seeded_inputs
(numpy legacy RandomState => reproducible without storing 2.3 MB of weights); a
checksum of both is stored so drift is detected.  Outputs stored per stage:
pre-U-Net planes, post-U-Net planes, the sampled 96-d features, the summed 32-d
`query_feature`, every head output, `infer_geo`.
"""
import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np

REF = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def import_reference():
    np.int = int  # noqa: binvox_rw.py:206
    import torch

    ts = types.ModuleType("torch_scatter")

    def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
        idx = index.expand_as(src)
        if out is None:
            shape = list(src.shape)
            shape[dim] = int(dim_size if dim_size is not None else idx.max() + 1)
            out = src.new_zeros(shape)
        out.scatter_add_(dim, idx, src)
        cnt = torch.zeros_like(out).scatter_add_(dim, idx, torch.ones_like(src)).clamp_(min=1)
        out.div_(cnt)
        return out

    ts.scatter_mean = scatter_mean
    ts.scatter_max = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError)
    sys.modules["torch_scatter"] = ts
    for m in ["matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d", "trimesh"]:
        s = MagicMock()
        s.__path__ = []
        sys.modules[m] = s
    sys.path.insert(0, REF)
    import vgn.networks as N  # noqa

    return N


def checksum(t):
    a = np.ascontiguousarray(t.detach().numpy()).astype(np.float64)
    return float(a.sum()), float(np.abs(a).sum())


def main():
    import torch
    from oracle import giga_oracle as O

    torch.set_num_threads(8)
    N = import_reference()
    out = {}
    for name in ["giga", "giga_aff", "giga_geo", "giga_detach"]:
        net = N.get_network(name)
        keys = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
        out[f"keys_{name}"] = np.array([k for k, _ in keys])
        out[f"shapes_{name}"] = np.array([str(s) for _, s in keys])
    net = N.get_network("giga")
    ref_keys = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
    assert sorted(ref_keys) == sorted(O.param_shapes()), "oracle.param_shapes() != reference state_dict"
    sd = O.seeded_state_dict(seed=1)
    net.load_state_dict(sd)
    net.eval()
    out["sd_checksum"] = np.array([checksum(torch.cat([v.flatten() for v in sd.values()]))])

    cases = {"a": dict(B=2, N=256, seed=0), "b": dict(B=1, N=64, seed=3)}
    with torch.no_grad():
        for tag, cfg in cases.items():
            x, p, pt = O.seeded_inputs(cfg["B"], cfg["N"], seed=cfg["seed"])
            out[f"{tag}_cfg"] = np.array([cfg["B"], cfg["N"], cfg["seed"]])
            out[f"{tag}_in_checksum"] = np.array([checksum(x), checksum(p), checksum(pt)])
            enc = net.encoder
            # stage 1: pre-U-Net planes: run encoder with the U-Net disabled
            unet = enc.unet
            enc.unet = None
            pre = enc(x)
            enc.unet = unet
            planes = net.encode_inputs(x)
            for k in ("xz", "xy", "yz"):
                if cfg["B"] == 1:   # full planes only for the single-scene case (fixture size)
                    out[f"{tag}_pre_{k}"] = pre[k].numpy()
                    out[f"{tag}_plane_{k}"] = planes[k].numpy()
                else:               # a strided sub-sample elsewhere
                    out[f"{tag}_pre_{k}"] = pre[k][:, ::4, ::3, ::3].numpy()
                    out[f"{tag}_plane_{k}"] = planes[k][:, ::4, ::3, ::3].numpy()
            dq = net.decoder_qual
            feat = torch.cat([dq.sample_plane_feature(p, planes[k], plane=k) for k in ("xz", "xy", "yz")], 1).transpose(1, 2)
            out[f"{tag}_feat96"] = feat.numpy()
            out[f"{tag}_qfeat32"] = net.query_feature(p, planes).numpy()
            qual, rot, width, tsdf = net(x, p, p_tsdf=pt)
            q2, r2, w2 = net(x, p)
            assert torch.equal(q2, qual) and torch.equal(r2, rot) and torch.equal(w2, width)
            out[f"{tag}_qual"], out[f"{tag}_rot"] = qual.numpy(), rot.numpy()
            out[f"{tag}_width"], out[f"{tag}_tsdf"] = width.numpy(), tsdf.numpy()
            out[f"{tag}_geo"] = net.infer_geo(x, pt).numpy()
            out[f"{tag}_occ_probs"] = net.decode_occ(pt, planes).probs.numpy()
    path = os.path.join(HERE, "giga_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
