#!/usr/bin/env python
"""Golden fixture for the VGN baseline network from the UNMODIFIED reference `ConvNet` (/root/reference/src/vgn/networks.py:48-63).

Run in the build container only (the GPU box has no /root/reference):   python tests/golden/make_vgn_golden.py
The reference module is imported with the shims of make_golden.py; parameters / inputs come from oracle.vgn_oracle's seeded generators
(a checksum of both is stored).  Stored for 2 scenes: the encoder output, the decoder output sub-sampled, the three output volumes at 8000
seeded voxel positions (`vox`) and float64 sums of the full volumes.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


def main():
    import torch
    from make_golden import import_reference
    import_reference()
    import vgn.networks as N
    from oracle import vgn_oracle as V

    torch.manual_seed(0)
    torch.set_num_threads(4)
    net = N.get_network("vgn")
    sd = V.seeded_state_dict(seed=3)
    net.load_state_dict(sd, strict=True)
    net.eval()
    x = V.seeded_inputs(2, seed=5)
    with torch.no_grad():
        e = net.encoder(x)
        d = net.decoder(e)
        qual, rot, width = net(x)
    chk = float(sum(v.double().abs().sum() for v in sd.values()) + x.double().sum())
    vox = np.sort(np.random.RandomState(11).choice(64000, 8000, replace=False))
    pick = lambda t: t.reshape(t.shape[0], t.shape[1], 64000)[:, :, vox].numpy().copy()
    out = dict(checksum=np.float64(chk), enc=e.numpy(), dec_s=d[:, :, ::4, ::4, ::4].numpy().copy(), vox=vox, qual=pick(qual), rot=pick(rot), width=pick(width),
               sums=np.array([qual.double().sum(), rot.double().abs().sum(), width.double().sum()]))
    np.savez_compressed(os.path.join(HERE, "vgn_golden.npz"), **out)
    print({k: getattr(v, "shape", v) for k, v in out.items()}, "qual>0.9:", int((qual > 0.9).sum()), "width range", float(width.min()), float(width.max()))


if __name__ == "__main__":
    main()
