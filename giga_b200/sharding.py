"""Scene sharding across the GPUs of one box.

Scenes are independent (no cross-batch op anywhere in the model), so the batch dimension
is split contiguously over ranks with NO data-path collective; the only exchange is the
final grasp-score reduction: every rank contributes, per scene, (best quality, arg-max
point index), 8 bytes/scene, gathered with one NCCL all-gather (two tensors).  The
per-scene arg-max kernel (giga_scene_argmax) writes straight into the rank's slice of the
gather buffer.  One process per GPU; `torch.distributed` is plumbing only.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_scenes: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced partition: the first (n_scenes % world) ranks get one extra scene."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, extra = divmod(n_scenes, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def max_shard(n_scenes: int, world: int) -> int:
    return -(-n_scenes // world)


def gather_scene_best(local_val: torch.Tensor, local_idx: torch.Tensor, n_scenes: int, group=None):
    """All-gather the per-scene (best value, arg-max index) of every rank's shard.

    local_val (n_local,) float32, local_idx (n_local,) int32 for the scenes shard_range() assigns to
    this rank.  Returns (val (n_scenes,), idx (n_scenes,)) identical on every rank.  Uneven shards are
    padded to the largest shard for the collective and trimmed afterwards.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    s, e = shard_range(n_scenes, rank, world)
    if local_val.numel() != e - s or local_idx.numel() != e - s:
        raise ValueError(f"rank {rank}: expected {e - s} scenes, got {local_val.numel()}")
    if world == 1:
        return local_val.clone(), local_idx.clone()
    m = max_shard(n_scenes, world)
    buf_v = torch.zeros((world, m), dtype=torch.float32, device=local_val.device)
    buf_i = torch.zeros((world, m), dtype=torch.int32, device=local_val.device)
    send_v = torch.zeros(m, dtype=torch.float32, device=local_val.device)
    send_i = torch.zeros(m, dtype=torch.int32, device=local_val.device)
    send_v[: e - s] = local_val
    send_i[: e - s] = local_idx
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(buf_v, send_v, group=group)
        dist.all_gather_into_tensor(buf_i, send_i, group=group)
    else:  # gloo (CPU tests): list form
        dist.all_gather(list(buf_v.unbind(0)), send_v, group=group)
        dist.all_gather(list(buf_i.unbind(0)), send_i, group=group)
    vals, idxs = [], []
    for r in range(world):
        rs, re = shard_range(n_scenes, r, world)
        vals.append(buf_v[r, : re - rs])
        idxs.append(buf_i[r, : re - rs])
    return torch.cat(vals), torch.cat(idxs)


def sharded_best_grasp(score_fn: Callable[[torch.Tensor, torch.Tensor], Tuple[torch.Tensor, torch.Tensor]],
                       tsdf: torch.Tensor, points: torch.Tensor, group=None):
    """Run `score_fn` on this rank's contiguous shard of the global batch and gather the result.

    score_fn(tsdf_shard (n,40,40,40), points_shard (n,N,3)) -> (best_val (n,), best_idx (n,) int32);
    in production it is `GigaScorer(net)` below (CUDA), in the CPU tests an oracle stand-in.
    `tsdf` / `points` are the GLOBAL batch (every rank holds or can index it); only the shard is touched.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = tsdf.shape[0]
    s, e = shard_range(n, rank, world)
    if e > s:
        v, i = score_fn(tsdf[s:e], points[s:e])
    else:
        v = torch.empty(0, dtype=torch.float32, device=tsdf.device)
        i = torch.empty(0, dtype=torch.int32, device=tsdf.device)
    return gather_scene_best(v.float(), i.to(torch.int32), n, group=group)


class SceneBestBuffer:
    """The all-gather buffer of the final grasp-score reduction, packed so that ONE collective moves both fields:
    buf[world][2][m] float32 with row 0 = best quality and row 1 = the arg-max index (int32 bits).  `val` / `idx` are
    this rank's contiguous slices of the CURRENT buffer -- giga_scene_argmax (or giga_forward's fused arg-max) writes
    straight into them -- and the exchange is a single in-place NCCL all-gather (8 bytes per scene).

    gather()        in-stream: the collective is enqueued behind the kernels on the current stream (simple; at 8 GPUs it put
                    ~20 us of launch + NVLink latency on every step's critical path, SCALE_r01).
    gather_async()  the collective runs on a side stream behind an event and the buffer ring advances (`depth` buffers): the next
                    step's kernels start at once and write the next buffer; wait(ticket) makes the current stream wait for a
                    gather and returns its (val, idx).  The latency-bound 8 B/scene exchange then overlaps the next step."""

    def __init__(self, n_local_max: int, device, group=None, depth: int = 2):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.bufs = [torch.zeros((self.world, 2, n_local_max), dtype=torch.float32, device=device) for _ in range(max(1, depth))]
        self.cur = 0
        dev = torch.device(device)
        self._cuda = dev.type == "cuda"
        self._side = torch.cuda.Stream(device=dev) if (self._cuda and self.world > 1) else None
        self._done = [None] * len(self.bufs)     # event: the gather that last used buffer i has finished

    @property
    def buf(self) -> torch.Tensor:
        return self.bufs[self.cur]

    @property
    def val(self) -> torch.Tensor:
        return self.buf[self.rank, 0]

    @property
    def idx(self) -> torch.Tensor:
        return self.buf[self.rank, 1].view(torch.int32)

    def _all_gather(self, buf):
        if dist.get_backend(self.group) == "nccl":
            dist.all_gather_into_tensor(buf, buf[self.rank], group=self.group)   # in place: send = own slice
        else:  # gloo (CPU tests): list form, no aliasing
            dist.all_gather(list(buf.unbind(0)), buf[self.rank].clone(), group=self.group)

    def gather(self):
        """-> (val (world, m) float32, idx (world, m) int32), identical on every rank (in-stream)."""
        if self.world > 1:
            self._all_gather(self.buf)
        return self.buf[:, 0], self.buf[:, 1].view(torch.int32)

    def gather_async(self) -> int:
        """Start the exchange of the current buffer off the critical path and advance to the next buffer; returns a ticket."""
        t = self.cur
        if self.world > 1:
            if self._side is not None:
                ready = torch.cuda.Event()
                ready.record()                                   # this step's arg-max has written the slice
                self._side.wait_event(ready)
                with torch.cuda.stream(self._side):
                    self._all_gather(self.bufs[t])
                    done = torch.cuda.Event()
                    done.record()
                self._done[t] = done
            else:
                self._all_gather(self.bufs[t])
        self.cur = (t + 1) % len(self.bufs)
        if self._done[self.cur] is not None:                     # the buffer we are about to overwrite: its last gather (depth steps ago)
            torch.cuda.current_stream().wait_event(self._done[self.cur])
        return t

    def wait(self, ticket: int):
        """Make the current stream wait for gather `ticket`; -> (val (world, m), idx (world, m)) of that exchange."""
        if self._done[ticket] is not None:
            torch.cuda.current_stream().wait_event(self._done[ticket])
        b = self.bufs[ticket]
        return b[:, 0], b[:, 1].view(torch.int32)

    def synchronize(self):
        if self._side is not None:
            self._side.synchronize()


class GigaScorer:
    """score_fn for the CUDA model: forward of the grasp heads + the fused per-scene arg-max kernel."""

    def __init__(self, net):
        self.net = net

    def __call__(self, tsdf: torch.Tensor, points: torch.Tensor):
        with torch.no_grad():
            c = self.net.encode_inputs(tsdf)
            qual, _, _ = self.net.decode(points, c)
            return self.net.scene_argmax(qual)
