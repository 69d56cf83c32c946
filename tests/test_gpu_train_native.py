"""GPU: the native training step (csrc/train_bwd.cuh, giga_train_forward / giga_train_backward) against CPU autograd through the
oracle (which is pinned to the unmodified reference network, tests/test_oracle_golden.py): outputs, every parameter gradient, the
model variants, both gradient-accumulation modes, and a few optimizer steps of a train_giga.py-style loop."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import giga_b200
from giga_b200 import training
from oracle import giga_oracle as O
from tests.util import make_net

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GRAD_TOL = 2e-4        # max |g - ref| / max |ref| per tensor (fp32 sums in a different order; atomics)


def _batch(B, No, seed, Ng=1):
    x, p, _ = O.seeded_inputs(B, Ng, seed=seed)
    _, _, pt = O.seeded_inputs(B, No, seed=seed + 1)
    g = torch.Generator().manual_seed(seed)
    label = (torch.rand(B, generator=g) > 0.4).float()
    rot_t = F.normalize(torch.randn(B, 2, 4, generator=g), dim=-1)
    width_t = torch.rand(B, generator=g) * 0.08
    occ_t = (torch.rand(B, No, generator=g) > 0.5).float()
    return x, p, pt, (label, rot_t, width_t, occ_t)


def _loss(out, y, dev):
    """train_giga.py:153-195 (select + loss_fn) in plain torch ops"""
    label, rot_t, width_t, occ_t = (t.to(dev) for t in y)
    qual, rot, width, occ = out
    qual, rot, width, occ = qual.squeeze(-1), rot.squeeze(1), width.squeeze(-1), torch.sigmoid(occ)
    l_qual = F.binary_cross_entropy(qual, label, reduction="none")
    l0 = 1.0 - (rot * rot_t[:, 0]).sum(-1).abs()
    l1 = 1.0 - (rot * rot_t[:, 1]).sum(-1).abs()
    l_rot = torch.min(l0, l1)
    l_width = F.mse_loss(40 * width, 40 * width_t, reduction="none")
    l_occ = F.binary_cross_entropy(occ, occ_t, reduction="none").mean(-1)
    return (l_qual + label * (l_rot + 0.01 * l_width) + l_occ).mean()


def _compare_grads(net, ref_leaves, second=None):
    """every parameter gradient against CPU autograd through the oracle.  The gradient of a ReLU / max-pool network is DISCONTINUOUS
    where a pre-activation is ~0 or two window entries nearly tie, and two correct fp32 forwards can land on different sides (measured:
    ATen's own GPU kernels differ from ATen's CPU kernels by 2.7e-3 on the B = 9 batch below, on exactly the tensors where this library
    does, while this library and ATen-on-GPU agree to 2e-5: profiles/r02s_grad_report_b9.txt).  So a tensor that misses GRAD_TOL against
    the CPU oracle must stay within 5e-2 of it AND match `second` -- the same function differentiated by PyTorch on the GPU -- to GRAD_TOL."""
    worst = ("", 0.0)
    for k, prm in net.named_parameters():
        r = ref_leaves[k].grad
        if r is None:
            assert prm.grad is None or float(prm.grad.abs().max()) == 0.0, k
            continue
        assert prm.grad is not None, k
        g = prm.grad.detach().cpu()
        err = ((g - r).abs().max() / (r.abs().max() + 1e-12)).item()
        if err >= GRAD_TOL and second is not None:
            assert err < 5e-2, (k, err)       # sanity net only: which near-ties flip depends on the host CPU's own rounding as well
            if callable(second):              # evaluated only when a tensor needs it
                second = second()
            err = ((g - second[k]).abs().max() / (r.abs().max() + 1e-12)).item()
        if err > worst[1]:
            worst = (k, err)
    assert worst[1] < GRAD_TOL, worst
    return worst


def _gpu_autograd_reference(sd, x, p, pt, y, name="giga"):
    """the same loss differentiated by PyTorch's own GPU kernels (fp32, TF32 off) through tests/torch_bridge.py"""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from tests.torch_bridge import bridged_forward
    net2 = make_net(name, sd, frozen=False)
    _loss(bridged_forward(net2, x.to(DEV), p.to(DEV), pt.to(DEV)), y, DEV).backward()
    return {k: v.grad.detach().cpu() for k, v in net2.named_parameters()}


@pytest.mark.parametrize("B,No,fwd_impl", [(4, 128, 1), (9, 300, 1), (4, 128, 0), (9, 300, 0)])
def test_native_forward_and_every_parameter_gradient(oracle_sd, B, No, fwd_impl):
    """B = 4 runs conv_in's fine tiling, B = 9 the 5-row tiling; No = 300 has a ragged last point tile; fwd_impl: the training forward on
    the tcgen05 kernels (device-packed operands, default) or on the fp32 FMA-pipe kernels"""
    net = make_net("giga", oracle_sd, frozen=False)
    net._engine_raw().set_option("train_forward_impl", fwd_impl)
    x, p, pt, y = _batch(B, No, seed=50 + B)
    out = net(x.to(DEV), p.to(DEV), p_tsdf=pt.to(DEV))
    assert all(o.requires_grad for o in out)                    # an ordinary differentiable module: no opt-in
    ref_leaves = {k: v.clone().requires_grad_(True) for k, v in oracle_sd.items()}
    ref_out = O.forward(ref_leaves, x, p, pt)
    for a, b, nm in zip(out, ref_out, ("qual", "rot", "width", "occ")):
        assert a.shape == b.shape
        assert (a.detach().cpu() - b.detach()).abs().max().item() <= 1e-4, nm
    n0 = net.gpu_launches
    _loss(out, y, DEV).backward()
    assert net.gpu_launches - n0 >= 30                          # the backward ran on this library's kernels
    _loss(ref_out, y, "cpu").backward()
    _compare_grads(net, ref_leaves, second=_gpu_autograd_reference(oracle_sd, x, p, pt, y))


@pytest.mark.parametrize("B,Ng,No", [(2, 130, 64), (1, 257, 1)])
def test_many_grasp_points_per_scene(oracle_sd, B, Ng, No):
    """grasp heads differentiated at more than one 128-point tile per scene (ragged last tile), a single scene, a single occupancy point"""
    net = make_net("giga", oracle_sd, frozen=False)
    x, p, _ = O.seeded_inputs(B, Ng, seed=61)
    _, _, pt = O.seeded_inputs(B, No, seed=62)
    g = torch.Generator().manual_seed(5)
    wq, wr, ww, wo = torch.randn(B, Ng, generator=g), torch.randn(B, Ng, 4, generator=g), torch.randn(B, Ng, generator=g), torch.randn(B, No, generator=g)
    lf = lambda o, dev: (o[0] * wq.to(dev)).sum() + (o[1] * wr.to(dev)).sum() + (o[2] * ww.to(dev)).sum() + (o[3] * wo.to(dev)).sum()
    out = net(x.to(DEV), p.to(DEV), p_tsdf=pt.to(DEV))
    leaves = {k: v.clone().requires_grad_(True) for k, v in oracle_sd.items()}
    ref = O.forward(leaves, x, p, pt)
    for a, b in zip(out, ref):
        assert (a.detach().cpu() - b.detach()).abs().max().item() <= 1e-4
    lf(out, DEV).backward()
    lf(ref, "cpu").backward()
    from tests.torch_bridge import bridged_forward
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    net2 = make_net("giga", oracle_sd, frozen=False)
    lf(bridged_forward(net2, x.to(DEV), p.to(DEV), pt.to(DEV)), DEV).backward()
    _compare_grads(net, leaves, second={k: v.grad.detach().cpu() for k, v in net2.named_parameters()})


def test_gradients_accumulate_and_upstream_scale(oracle_sd):
    """a second backward adds to .grad (autograd semantics); gradients are linear in the upstream gradient"""
    net = make_net("giga", oracle_sd, frozen=False)
    x, p, pt, y = _batch(4, 64, seed=7)
    xd, pd, ptd = x.to(DEV), p.to(DEV), pt.to(DEV)
    _loss(net(xd, pd, p_tsdf=ptd), y, DEV).backward()
    g1 = {k: v.grad.clone() for k, v in net.named_parameters()}
    (2.0 * _loss(net(xd, pd, p_tsdf=ptd), y, DEV)).backward()
    for k, v in net.named_parameters():
        # the sums run through floating-point atomics: two evaluations differ at the 1e-6 level of the tensor's largest entry
        assert float((v.grad - 3.0 * g1[k]).abs().max()) <= 2e-5 * 3.0 * float(g1[k].abs().max()) + 1e-12, k


@pytest.mark.parametrize("name", ["giga_aff", "giga_geo", "giga_detach"])
def test_model_variants(oracle_sd, name):
    net = make_net(name, oracle_sd, frozen=False)
    x, p, pt, y = _batch(4, 96, seed=21)
    sd = {k: v for k, v in oracle_sd.items() if k in net.state_dict()}
    ref_leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    label, rot_t, width_t, occ_t = y
    if name == "giga_aff":
        out = net(x.to(DEV), p.to(DEV))
        ref = O.forward(ref_leaves, x, p, None)
        lf = lambda o, dev: (o[0].sum() + (o[1] * rot_t[:, :1].to(dev)).sum() + 3.0 * o[2].sum())
    elif name == "giga_geo":
        out = (net(x.to(DEV), pt.to(DEV), pt.to(DEV)),)
        ref = (O.infer_geo(ref_leaves, x, pt),)
        lf = lambda o, dev: F.binary_cross_entropy_with_logits(o[0], occ_t.to(dev))
    else:
        out = net(x.to(DEV), p.to(DEV), p_tsdf=pt.to(DEV))
        ref = O.forward(ref_leaves, x, p, pt, detach_tsdf=True)
        lf = lambda o, dev: _loss(o, y, dev)
    for a, b in zip(out, ref):
        assert (a.detach().cpu() - b.detach()).abs().max().item() <= 1e-4
    lf(out, DEV).backward()
    lf(ref, "cpu").backward()

    def gpu_autograd():
        from tests.torch_bridge import bridged_forward
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        net2 = make_net(name, oracle_sd, frozen=False)
        xd, pd, ptd = x.to(DEV), p.to(DEV), pt.to(DEV)
        out2 = bridged_forward(net2, xd, pd, None) if name == "giga_aff" else bridged_forward(net2, xd, None if name == "giga_geo" else pd, ptd)
        lf(out2, DEV).backward()
        return {k: (v.grad.detach().cpu() if v.grad is not None else torch.zeros_like(v).cpu()) for k, v in net2.named_parameters()}

    _compare_grads(net, ref_leaves, second=gpu_autograd)


def test_position_gradients_match_autograd(oracle_sd):
    """d(outputs)/d(p) from the backward kernels (fc_p + grid_sampler's grid gradient, one-sided clamps, border clip) against CPU autograd
    through the oracle, together with the parameter gradients of the same backward; points on / beyond the cube faces included"""
    net = make_net("giga", oracle_sd, frozen=False)
    x, p, _ = O.seeded_inputs(3, 200, seed=17)          # edge cases on: points outside [-0.5, 0.5] and exactly on +-0.5
    pd = p.to(DEV).requires_grad_(True)
    q, r, w = net(x.to(DEV), pd)
    wts = torch.linspace(0.5, 1.5, 200)
    (q * wts.to(DEV)).sum().backward(retain_graph=False)
    leaves = {k: v.clone().requires_grad_(True) for k, v in oracle_sd.items()}
    pr = p.clone().requires_grad_(True)
    qr, rr, wr = O.forward(leaves, x, pr)
    (qr * wts).sum().backward()
    scale = float(pr.grad.abs().max())
    assert scale > 0
    assert float((pd.grad.cpu() - pr.grad).abs().max()) <= 2e-4 * scale
    _compare_grads(net, {k: v for k, v in leaves.items()})


def test_one_forward_at_a_time(oracle_sd):
    net = make_net("giga", oracle_sd, frozen=False)
    x, p, pt, y = _batch(2, 32, seed=3)
    xd, pd, ptd = x.to(DEV), p.to(DEV), pt.to(DEV)
    out1 = net(xd, pd, p_tsdf=ptd)
    out2 = net(xd, pd, p_tsdf=ptd)            # overwrites the kept activations
    with pytest.raises(giga_b200.GigaError):
        _loss(out1, y, DEV).backward()
    _loss(out2, y, DEV).backward()
    # inference under no_grad between a forward and its backward is refused the same way only if it is a training forward:
    with torch.no_grad():
        net(xd, pd, p_tsdf=ptd)


def test_training_loop_matches_oracle_training(oracle_sd):
    """train_giga.py-style steps: native forward/backward + fused loss + flat Adam (direct gradient accumulation) against
    torch.optim.Adam over CPU autograd through the oracle; the parameters after 3 steps and the loss trajectory agree"""
    net = make_net("giga", oracle_sd, frozen=False)
    opt = training.Adam(net.parameters(), lr=2e-4)
    ref = {k: v.clone().requires_grad_(True) for k, v in oracle_sd.items()}
    ropt = torch.optim.Adam(list(ref.values()), lr=2e-4)
    losses, rlosses = [], []
    for step in range(3):
        x, p, pt, y = _batch(8, 256, seed=100 + step)
        opt.zero_grad()
        out = net(x.to(DEV), p.to(DEV), p_tsdf=pt.to(DEV))
        loss, _ = training.loss_fn(training.select(out), tuple(t.to(DEV) for t in y))
        loss.backward()
        opt.step()
        losses.append(float(loss))
        ropt.zero_grad()
        rl = _loss(O.forward(ref, x, p, pt), y, "cpu")
        rl.backward()
        ropt.step()
        rlosses.append(float(rl))
    np.testing.assert_allclose(losses, rlosses, rtol=2e-4)
    bad = total = 0
    for k, v in net.named_parameters():
        # Adam's first steps move every weight by ~lr whatever the gradient's size (an element whose gradient is at the noise level of
        # the fp32 sums can move the other way): compare the displacements and allow a vanishing fraction of such elements
        d = (v.detach().cpu() - oracle_sd[k]), (ref[k].detach() - oracle_sd[k])
        bad += int(((d[0] - d[1]).abs() > 0.05 * 3 * 2e-4).sum())
        total += v.numel()
    assert bad <= 5e-3 * total, (bad, total)
    # and the inference path picks the trained weights up (the optimizer bumped the version counters: re-commit)
    trained = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    with torch.no_grad():
        x, p, pt, _ = _batch(2, 64, seed=5)
        got = net(x.to(DEV), p.to(DEV), p_tsdf=pt.to(DEV))
        want = O.forward(trained, x, p, pt)
    for a, b in zip(got, want):
        assert (a.cpu() - b).abs().max().item() <= 1e-4
