"""GPU: the fused loss kernel and the flat Adam kernel (csrc/train.cuh) against the oracle and the reference-made fixture."""
import os

import numpy as np
import pytest
import torch

import giga_b200
from giga_b200 import training
from oracle import train_oracle as T

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "train_golden.npz")


@pytest.mark.parametrize("tag,B,M,seed", [("a", 32, 2048, 0), ("b", 5, 7, 1)])
def test_fused_loss_value_and_gradients(tag, B, M, seed):
    gold = np.load(GOLD)
    arrs = [torch.from_numpy(a).cuda() for a in T.seeded_batch(B, M, seed)]
    preds = [a.clone().requires_grad_(True) for a in arrs[:4]]
    loss, d = training.loss_fn(tuple(preds), tuple(arrs[4:]))
    loss.backward()
    means = np.array([d[k].item() for k in ("loss_qual", "loss_rot", "loss_width", "loss_occ", "loss_all")])
    assert float(loss) == means[4]
    np.testing.assert_allclose(means, gold[f"{tag}_means"], rtol=2e-6)          # fp32 sums in a different order than ATen's
    omeans, ograds = T.loss(*[a.cpu().numpy() for a in arrs])
    np.testing.assert_allclose(means, omeans, rtol=2e-6)
    for name, p, og in zip(("label", "rot", "width", "occ"), preds, ograds):
        ref = gold[f"{tag}_g_{name}"]
        got = p.grad.cpu().numpy()
        scale = max(1.0, np.abs(ref).max())
        np.testing.assert_allclose(got, ref, rtol=2e-5, atol=1e-9 * scale, err_msg=name)
        np.testing.assert_allclose(got, og, rtol=2e-5, atol=1e-9 * scale, err_msg=name)


def test_fused_loss_scales_with_upstream_gradient_and_is_deterministic():
    arrs = [torch.from_numpy(a).cuda() for a in T.seeded_batch(32, 512, 4)]
    out = []
    for k in (1.0, 3.0):
        preds = [a.clone().requires_grad_(True) for a in arrs[:4]]
        loss, _ = training.loss_fn(tuple(preds), tuple(arrs[4:]))
        (loss * k).backward()
        out.append([float(loss)] + [p.grad.clone() for p in preds])
    assert out[0][0] == out[1][0]
    for a, b in zip(out[0][1:], out[1][1:]):
        torch.testing.assert_close(a * 3.0, b, rtol=1e-6, atol=0)
    with pytest.raises(giga_b200.GigaError):
        training.loss_fn((arrs[0], arrs[1][:, :3], arrs[2], arrs[3]), tuple(arrs[4:]))


@pytest.mark.parametrize("tag,kw", [("plain", dict(lr=2e-4)), ("wd", dict(lr=1e-2, betas=(0.8, 0.95), eps=1e-6, weight_decay=0.1))])
def test_flat_adam_matches_torch_optim(tag, kw):
    gold = np.load(GOLD)
    p0, gs = gold["adam_p0"], gold["adam_g"]
    # three parameters of ragged sizes (padding of the flat layout), one launch per step
    cuts = [0, 7, 500, 1003]
    params = [torch.nn.Parameter(torch.from_numpy(p0[a:b].copy()).cuda()) for a, b in zip(cuts[:-1], cuts[1:])]
    opt = training.Adam(params, **kw)
    eng = training._engine(params[0].device)
    n0 = eng.launches
    want = {0: 0, 5: 1, 11: 2}
    for step, g in enumerate(gs):
        opt.zero_grad()
        for p, (a, b) in zip(params, zip(cuts[:-1], cuts[1:])):
            p.grad.add_(torch.from_numpy(g[a:b].copy()).cuda())      # what backward() does: accumulate into the flat views
        opt.step()
        if step in want:
            got = torch.cat([p.detach().reshape(-1) for p in params]).cpu().numpy()
            np.testing.assert_allclose(got, gold[f"adam_{tag}"][want[step]], rtol=2e-5, atol=2e-6)
    assert eng.launches - n0 == len(gs)
    traj = T.adam(p0, gs, **kw)
    np.testing.assert_allclose(torch.cat([p.detach().reshape(-1) for p in params]).cpu().numpy(), traj[-1], rtol=2e-5, atol=2e-6)


def test_adam_updates_the_model_in_place():
    """the re-pointed parameters stay the module's parameters: a GIGA forward after step() runs with the new values"""
    from oracle import giga_oracle as O
    sd = O.seeded_state_dict(seed=1)
    net = giga_b200.get_network("giga")
    net.load_state_dict(sd)
    net = net.cuda()
    opt = training.Adam(net.parameters(), lr=1e-2)
    x, p, pt = O.seeded_inputs(1, 64, seed=2)
    with torch.no_grad():
        before = net(x.cuda(), p.cuda(), p_tsdf=pt.cuda())[0].clone()
    opt.zero_grad()
    for q in net.parameters():
        q.grad.fill_(1.0)
    opt.step()
    new_sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    assert all(torch.allclose(new_sd[k], sd[k] - 1e-2, atol=1e-6) for k in sd)      # first Adam step = -lr * sign(g)
    with torch.no_grad():
        after = net(x.cuda(), p.cuda(), p_tsdf=pt.cuda())
    ref = O.forward(new_sd, x, p, pt)
    assert (after[0].cpu() - ref[0]).abs().max().item() < 1e-4
    assert not torch.equal(after[0], before)
