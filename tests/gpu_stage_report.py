#!/usr/bin/env python
"""GPU bring-up aid: compare EVERY stage of the CUDA path with the CPU oracle and print a table
(does not stop at the first mismatch).  Run on the GPU box:  python tests/gpu_stage_report.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import giga_b200  # noqa: E402
from oracle import giga_oracle as O  # noqa: E402


def main(B=2, N=300):
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    sd = O.seeded_state_dict(seed=1)
    net = giga_b200.get_network("giga")
    net.load_state_dict(sd)
    net = net.to(dev)
    x, p, pt = O.seeded_inputs(B, N, seed=0)
    rows = []

    def cmp(name, got, ref):
        got = got.detach().float().cpu()
        d = (got - ref).abs()
        rows.append((name, tuple(ref.shape), float(d.max()), float(d.mean()), float(ref.abs().max()), bool(torch.isfinite(got).all())))

    with torch.no_grad():
        pre = O.plane_features_pre_unet(sd, x)
        pre_stack = torch.stack([pre[k] for k in O.PLANES])  # [3][B][32][40][40]
        cap = {}
        ref_planes = torch.stack([O.unet_forward(sd, pre_stack.reshape(3 * B, 32, 40, 40), capture=cap).reshape(3, B, 32, 40, 40)[i]
                                  for i in range(3)])
        for eimpl, etag in ((0, "ffma"), (1, "tc")):
            net._engine().set_option("encoder_impl", eimpl)
            c = net.encode_inputs(x.to(dev))
            torch.cuda.synchronize()
            cmp(etag + ".pre", net.debug_activation("pre", B), pre_stack)
            for k in ["d0c1", "d0c2", "p0", "d1c1", "d1c2", "p1", "d2c1", "d2c2", "u0", "u0c1", "u0c2", "u1", "u1c1", "u1c2"]:
                try:
                    got = net.debug_activation(k, B)
                except giga_b200.GigaError:
                    continue
                cmp(etag + "." + k, got, cap[k].reshape(got.shape))
            cmp(etag + ".planes", torch.stack([c[k] for k in O.PLANES]), ref_planes)
        planes_ref = {k: ref_planes[i] for i, k in enumerate(O.PLANES)}
        cmp("feat96", net.sample_feature(p.to(dev), c, "concat"), O.sample_concat_feature(p, planes_ref))
        cmp("qfeat32", net.query_feature(p.to(dev), c), O.query_feature(p, planes_ref))
        rq, rr, rw, ro = O.forward(sd, x, p, pt)
        for impl, tag in ((0, "ffma"), (1, "tc")):
            net._engine().set_option("decoder_impl", impl)
            qual, rot, width, occ = net(x.to(dev), p.to(dev), p_tsdf=pt.to(dev))
            torch.cuda.synchronize()
            cmp(tag + ".qual", qual, rq); cmp(tag + ".rot", rot, rr); cmp(tag + ".width", width, rw); cmp(tag + ".occ", occ, ro)
        hq = net.forward_host(x.pin_memory(), p.pin_memory(), pt.pin_memory())
        cmp("host.qual", hq[0], rq); cmp("host.rot", hq[1], rr); cmp("host.width", hq[2], rw); cmp("host.occ", hq[3], ro)
        bv, bi = net.scene_argmax(qual)
        rows.append(("argmax", (B,), float((bi.cpu().long() - rq.argmax(1)).abs().max()), 0.0, 0.0, True))
    print(f"{'stage':<14}{'shape':<24}{'max|d|':>12}{'mean|d|':>12}{'max|ref|':>12}  finite")
    for r in rows:
        print(f"{r[0]:<14}{str(r[1]):<24}{r[2]:>12.3e}{r[3]:>12.3e}{r[4]:>12.3e}  {r[5]}")
    print("launches:", net.gpu_launches)


if __name__ == "__main__":
    main()
