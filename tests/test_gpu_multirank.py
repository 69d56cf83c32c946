"""Multi-rank GPU test (NCCL): scenes sharded over 2 ranks, the per-scene (best quality, arg-max) written by the fused arg-max kernel into
the packed gather buffer and exchanged with the in-place NCCL all-gather -- in-stream and pipelined on the side stream -- must equal the
unsharded result bit for bit.  Needs >= 2 visible GPUs (run under `gpurun --gpus 2`); skipped otherwise."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_scenes, out):
    import torch.distributed as dist

    from giga_b200 import sharding
    from oracle import giga_oracle as O
    from tests.util import make_net

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        net = make_net("giga", device=dev)
        x, p, pt = O.seeded_inputs(n_scenes, 300, seed=3)
        xd, pd, ptd = x.to(dev), p.to(dev), pt.to(dev)
        ok = True
        with torch.no_grad():
            full_q = net(xd, pd)[0]
            ref_v, ref_i = net.scene_argmax(full_q)                       # unsharded, this rank's GPU
            s0, e0 = sharding.shard_range(n_scenes, rank, world)
            m = sharding.max_shard(n_scenes, world)
            best = sharding.SceneBestBuffer(m, dev, depth=2)
            tickets = []
            for step in range(4):                                          # ring of two buffers, exchanges overlap the next step
                net.forward_with_argmax(xd[s0:e0], pd[s0:e0], ptd[s0:e0], best.val[: e0 - s0], best.idx[: e0 - s0])
                tickets.append(best.gather_async())
                if step >= 1:
                    gv, gi = best.wait(tickets[step - 1])
                    torch.cuda.current_stream().synchronize()
                    for r in range(world):
                        rs, re = sharding.shard_range(n_scenes, r, world)
                        ok = ok and torch.equal(gv[r, : re - rs], ref_v[rs:re]) and torch.equal(gi[r, : re - rs], ref_i[rs:re])
            best.synchronize()
            v, i = sharding.sharded_best_grasp(sharding.GigaScorer(net), xd, pd)   # in-stream form
            ok = ok and torch.equal(v, ref_v) and torch.equal(i, ref_i)
        out[rank] = int(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_scenes", [6, 5])
def test_sharded_argmax_allgather_equals_unsharded_nccl(n_scenes):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    port = _free_port()
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, n_scenes, out), nprocs=2, join=True)
    assert [out[r] for r in range(2)] == [1, 1]


def _train_worker(rank, world, port, out):
    """data-parallel training step: each rank differentiates its shard of the batch with the native backward, the flat gradient buffer is
    all-reduced (sum / world) over NCCL in place, Adam steps -- gradients and updated parameters must equal the single-GPU step on the
    whole batch (DDP grads == single-GPU grads of the global batch)."""
    import torch.distributed as dist

    from giga_b200 import training
    from oracle import giga_oracle as O
    from tests.test_gpu_train_native import _batch
    from tests.util import make_net

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        Bg = 8
        x, p, pt, y = _batch(Bg, 200, seed=31)
        to = lambda t: t.to(dev)

        def grads_and_params(sl, allreduce):
            net = make_net("giga", device=dev, frozen=False)
            opt = training.Adam(net.parameters(), lr=1e-3)
            opt.zero_grad()
            out_ = net(to(x[sl]), to(p[sl]), p_tsdf=to(pt[sl]))
            loss, _ = training.loss_fn(training.select(out_), tuple(to(t[sl]) for t in y))
            loss.backward()
            if allreduce:
                opt.allreduce_gradients()
            g = torch.cat([q.grad.reshape(-1) for q in net.parameters()]).clone()
            opt.step()
            return g, torch.cat([q.detach().reshape(-1) for q in net.parameters()]).clone()

        per = Bg // world
        g_ddp, p_ddp = grads_and_params(slice(rank * per, (rank + 1) * per), True)
        g_one, p_one = grads_and_params(slice(0, Bg), False)
        scale = float(g_one.abs().max())
        ok = float((g_ddp - g_one).abs().max()) <= 2e-4 * scale
        # every rank holds the same reduced gradients and parameters
        gathered = [torch.empty_like(p_ddp) for _ in range(world)]
        dist.all_gather(gathered, p_ddp)
        ok = ok and all(torch.equal(gathered[0], t) for t in gathered)
        ok = ok and float((p_ddp - p_one).abs().max()) <= 2.1e-3       # one Adam step moves a weight by at most lr (sign flips at the noise level: 2 lr)
        ok = ok and float(((p_ddp - p_one).abs() > 1e-4).float().mean()) < 1e-3
        out[rank] = int(ok)
    finally:
        dist.destroy_process_group()


def test_data_parallel_training_step_equals_single_gpu_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    port = _free_port()
    out = mp.Manager().dict()
    mp.spawn(_train_worker, args=(2, port, out), nprocs=2, join=True)
    assert [out[r] for r in range(2)] == [1, 1]
