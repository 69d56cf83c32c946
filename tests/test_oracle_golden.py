"""Pin the CPU oracle (oracle/giga_oracle.py) against fixtures produced by the
unmodified reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import giga_oracle as O

TOL = 2e-5  # same ATen ops as the reference on the same machine class; summation order only


def _chk(t):
    a = np.ascontiguousarray(t.numpy()).astype(np.float64)
    return [float(a.sum()), float(np.abs(a).sum())]


def _sub(t, full):
    return t if full else t[:, ::4, ::3, ::3]


def test_param_inventory_matches_reference(golden):
    keys = list(golden["keys_giga"])
    shapes = list(golden["shapes_giga"])
    ours = dict(O.param_shapes())
    assert len(keys) == len(ours)
    for k, s in zip(keys, shapes):
        assert str(tuple(ours[k])) == s, k
    assert sum(int(np.prod(s)) for s in ours.values()) == 581863  # SURVEY.md quick facts
    # giga_aff lacks decoder_tsdf; giga_geo has only encoder + decoder_tsdf
    assert sorted(golden["keys_giga_aff"]) == sorted(k for k, _ in O.param_shapes(with_tsdf=False))
    assert sorted(golden["keys_giga_geo"]) == sorted(k for k, _ in O.param_shapes(grasp_heads=False))
    assert sorted(golden["keys_giga_detach"]) == sorted(keys)


def test_seeded_streams_are_stable(golden, oracle_sd):
    flat = torch.cat([v.flatten() for v in oracle_sd.values()])
    np.testing.assert_allclose(_chk(flat), golden["sd_checksum"][0], rtol=1e-12)
    for tag in ("a", "b"):
        B, N, seed = (int(v) for v in golden[f"{tag}_cfg"])
        x, p, pt = O.seeded_inputs(B, N, seed=seed)
        np.testing.assert_allclose([_chk(x), _chk(p), _chk(pt)], golden[f"{tag}_in_checksum"], rtol=1e-12)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_oracle_stages_match_reference(golden, oracle_sd, tag):
    B, N, seed = (int(v) for v in golden[f"{tag}_cfg"])
    x, p, pt = O.seeded_inputs(B, N, seed=seed)
    sd = oracle_sd
    with torch.no_grad():
        pre = O.plane_features_pre_unet(sd, x)
        planes = O.encode_inputs(sd, x)
        for k in O.PLANES:
            np.testing.assert_allclose(_sub(pre[k], B == 1).numpy(), golden[f"{tag}_pre_{k}"], atol=TOL, rtol=0)
            np.testing.assert_allclose(_sub(planes[k], B == 1).numpy(), golden[f"{tag}_plane_{k}"], atol=TOL, rtol=0)
        np.testing.assert_allclose(O.sample_concat_feature(p, planes).numpy(), golden[f"{tag}_feat96"], atol=TOL, rtol=0)
        np.testing.assert_allclose(O.query_feature(p, planes).numpy(), golden[f"{tag}_qfeat32"], atol=TOL, rtol=0)
        qual, rot, width, tsdf = O.forward(sd, x, p, pt)
        np.testing.assert_allclose(qual.numpy(), golden[f"{tag}_qual"], atol=TOL, rtol=0)
        np.testing.assert_allclose(rot.numpy(), golden[f"{tag}_rot"], atol=TOL, rtol=0)
        np.testing.assert_allclose(width.numpy(), golden[f"{tag}_width"], atol=TOL, rtol=0)
        np.testing.assert_allclose(tsdf.numpy(), golden[f"{tag}_tsdf"], atol=TOL, rtol=0)
        np.testing.assert_allclose(O.infer_geo(sd, x, pt).numpy(), golden[f"{tag}_geo"], atol=TOL, rtol=0)
        np.testing.assert_allclose(torch.sigmoid(tsdf).numpy(), golden[f"{tag}_occ_probs"], atol=TOL, rtol=0)
        assert (qual.argmax(1).numpy() == golden[f"{tag}_qual"].argmax(1)).all()


def test_plane_mean_identity(oracle_sd):
    """with 40^3 voxels onto 40^2 cells every cell averages exactly the
    40 voxels along the perpendicular axis -- the fact the CUDA encoder relies on."""
    x, _, _ = O.seeded_inputs(1, 8, seed=5)
    with torch.no_grad():
        pre = O.plane_features_pre_unet(oracle_sd, x)
        f = torch.relu(torch.nn.functional.conv3d(x.unsqueeze(1), oracle_sd["encoder.conv_in.weight"],
                                                  oracle_sd["encoder.conv_in.bias"], padding=1))  # [b,c,ix,iy,iz]
    np.testing.assert_allclose(pre["xz"].numpy(), f.mean(3).permute(0, 1, 3, 2).numpy(), atol=2e-6)  # [c,iz,ix]
    np.testing.assert_allclose(pre["xy"].numpy(), f.mean(4).permute(0, 1, 3, 2).numpy(), atol=2e-6)  # [c,iy,ix]
    np.testing.assert_allclose(pre["yz"].numpy(), f.mean(2).permute(0, 1, 3, 2).numpy(), atol=2e-6)  # [c,iz,iy]
