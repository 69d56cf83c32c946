"""CPU: the loss / Adam oracle against the fixture made from the reference's own loss functions and torch.optim.Adam
(tests/golden/make_train_golden.py), and the host-side guards of giga_b200.training."""
import os

import numpy as np
import pytest
import torch

from oracle import train_oracle as T

GOLD = os.path.join(os.path.dirname(__file__), "golden", "train_golden.npz")


@pytest.mark.parametrize("tag,B,M,seed", [("a", 32, 2048, 0), ("b", 5, 7, 1)])
def test_loss_oracle_matches_reference(tag, B, M, seed):
    gold = np.load(GOLD)
    arrs = T.seeded_batch(B, M, seed)
    assert np.float64(sum(float(np.abs(a.astype(np.float64)).sum()) for a in arrs)) == gold[f"{tag}_checksum"]
    means, grads = T.loss(*arrs)
    np.testing.assert_allclose(means, gold[f"{tag}_means"], rtol=2e-6)
    for name, g in zip(("label", "rot", "width", "occ"), grads):
        ref = gold[f"{tag}_g_{name}"]
        np.testing.assert_allclose(g, ref, rtol=2e-5, atol=1e-9 * max(1.0, np.abs(ref).max()), err_msg=name)


@pytest.mark.parametrize("tag,kw", [("plain", dict(lr=2e-4)), ("wd", dict(lr=1e-2, betas=(0.8, 0.95), eps=1e-6, weight_decay=0.1))])
def test_adam_oracle_matches_torch_optim(tag, kw):
    gold = np.load(GOLD)
    traj = T.adam(gold["adam_p0"], gold["adam_g"], **kw)
    np.testing.assert_allclose(traj[[0, 5, 11]], gold[f"adam_{tag}"], rtol=1e-5, atol=1e-6)


def test_training_mirror_has_no_cpu_path():
    import giga_b200
    from giga_b200 import training
    arrs = [torch.from_numpy(a) for a in T.seeded_batch(4, 4, 0)]
    with pytest.raises(giga_b200.GigaError):
        training.loss_fn(tuple(arrs[:4]), tuple(arrs[4:]))
    with pytest.raises(giga_b200.GigaError):
        training.Adam([torch.nn.Parameter(torch.zeros(4))], lr=1e-3)
    with pytest.raises(ValueError):
        training.Adam([torch.nn.Parameter(torch.zeros(4))], lr=-1.0)
    q, r, w, o = training.select((torch.zeros(3, 1), torch.zeros(3, 1, 4), torch.zeros(3, 1), torch.zeros(3, 9)))   # train_giga.py:153-158
    assert q.shape == (3,) and r.shape == (3, 4) and w.shape == (3,) and float(o[0, 0]) == 0.5
