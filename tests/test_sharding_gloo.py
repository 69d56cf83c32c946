"""world_size-2 (and 3, uneven) gloo tests of the scene-sharding host logic on CPU.  The local
compute is an oracle stand-in (the product's CUDA scorer cannot run without a GPU); what is
under test is the partition / padding / all-gather / trim logic of giga_b200.sharding."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from giga_b200 import sharding


def test_shard_range_partitions():
    for n in (1, 2, 5, 32, 33, 256):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and b >= a
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and max(sizes) == sharding.max_shard(n, world)
    with pytest.raises(ValueError):
        sharding.shard_range(4, 4, 4)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _score(tsdf, points):
    # deterministic stand-in "model": quality = a fixed function of scene and point
    q = (points.sum(-1) * tsdf.flatten(1).mean(1, keepdim=True)).sin()
    v, i = q.max(1)
    return v, i.to(torch.int32)


def _worker(rank, world, port, n_scenes, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        tsdf = torch.rand(n_scenes, 40, 40, 40, generator=g)
        pts = torch.rand(n_scenes, 64, 3, generator=g) - 0.5
        v, i = sharding.sharded_best_grasp(_score, tsdf, pts)
        rv, ri = _score(tsdf, pts)
        ok = torch.equal(v, rv) and torch.equal(i, ri)
        # mismatched shard size is rejected
        try:
            sharding.gather_scene_best(torch.zeros(n_scenes + 1), torch.zeros(n_scenes + 1, dtype=torch.int32), n_scenes)
            ok = False
        except ValueError:
            pass
        # packed single-collective buffer (what bench.py / the serving loop use): every rank fills its own slice
        m = sharding.max_shard(n_scenes, world)
        best = sharding.SceneBestBuffer(m, "cpu")
        s0, e0 = sharding.shard_range(n_scenes, rank, world)
        best.val[: e0 - s0] = rv[s0:e0]
        best.idx[: e0 - s0] = ri[s0:e0]
        gv, gi = best.gather()
        for r in range(world):
            rs, re = sharding.shard_range(n_scenes, r, world)
            ok = ok and torch.equal(gv[r, : re - rs], rv[rs:re]) and torch.equal(gi[r, : re - rs], ri[rs:re])
        # pipelined form: three "steps" through a ring of two buffers; every ticket returns its own step's exchange
        ring = sharding.SceneBestBuffer(m, "cpu", depth=2)
        tickets = []
        for step in range(3):
            ring.val[: e0 - s0] = rv[s0:e0] + step
            ring.idx[: e0 - s0] = ri[s0:e0] + step
            tickets.append(ring.gather_async())
            if step >= 1:                                  # consume the previous step's exchange while "computing" this one
                pv, pi = ring.wait(tickets[step - 1])
                for r in range(world):
                    rs, re = sharding.shard_range(n_scenes, r, world)
                    ok = ok and torch.equal(pv[r, : re - rs], rv[rs:re] + (step - 1)) and torch.equal(pi[r, : re - rs], ri[rs:re] + (step - 1))
        ring.synchronize()
        out[rank] = int(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_scenes", [(2, 8), (2, 5), (3, 7), (2, 1)])
def test_sharded_equals_unsharded_gloo(world, n_scenes):
    port = _free_port()
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, n_scenes, out), nprocs=world, join=True)
    assert [out[r] for r in range(world)] == [1] * world


def test_single_process_path():
    g = torch.Generator().manual_seed(1)
    tsdf = torch.rand(3, 40, 40, 40, generator=g)
    pts = torch.rand(3, 16, 3, generator=g) - 0.5
    v, i = sharding.sharded_best_grasp(_score, tsdf, pts)
    rv, ri = _score(tsdf, pts)
    assert torch.equal(v, rv) and torch.equal(i, ri)


def _grad_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from giga_b200.training import allreduce_gradients
        torch.manual_seed(0)
        params = [torch.nn.Parameter(torch.randn(7, 3)), torch.nn.Parameter(torch.randn(5)), torch.nn.Parameter(torch.randn(2, 2))]
        params[0].grad = torch.full((7, 3), float(rank + 1))
        params[1].grad = torch.arange(5.0) * (rank + 1)
        # params[2] has no gradient on any rank -> zeros
        flat = allreduce_gradients(params)
        mean = sum(range(1, world + 1)) / world
        ok = torch.allclose(params[0].grad, torch.full((7, 3), mean)) and torch.allclose(params[1].grad, torch.arange(5.0) * mean)
        ok = ok and params[2].grad is not None and params[2].grad.abs().max().item() == 0 and flat.numel() == 21 + 5 + 4
        out[rank] = int(ok)
    finally:
        dist.destroy_process_group()


def test_flat_gradient_allreduce_gloo():
    port = _free_port()
    out = mp.Manager().dict()
    mp.spawn(_grad_worker, args=(2, port, out), nprocs=2, join=True)
    assert [out[r] for r in range(2)] == [1, 1]
    from giga_b200.training import allreduce_gradients
    assert allreduce_gradients([torch.nn.Parameter(torch.zeros(3))]) is None   # single process: no-op
