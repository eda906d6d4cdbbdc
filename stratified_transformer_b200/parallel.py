"""Data parallelism by scene: the only way this path shards.

Windows never cross scenes (the batch id is the slowest voxel coordinate) and FPS is per
scene, so scenes are independent units: rank g takes scenes {g, g + G, ...} exactly like the reference's
DistributedSampler (/root/reference/train.py:199-202) and there is NO collective on the data path.  The only
exchange is the training-time gradient all-reduce of the replicated parameters (reference: DDP, train.py:161),
done here as one flat NCCL all-reduce per step (a few MB: rel-pos tables + qkv/proj weights), which is noise next
to a step and therefore left to NCCL over NVLink/NVSwitch rather than a fused kernel.
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_scenes(n_scenes: int, rank: int, world_size: int) -> List[int]:
    """Scene ids owned by `rank` (round-robin, the DistributedSampler order without shuffling)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    return list(range(rank, n_scenes, world_size))


def shard_batch(xyz: torch.Tensor, offset: torch.Tensor, rank: int, world_size: int, *per_point: torch.Tensor):
    """Cut a collated batch (xyz [N,3], cumulative offset [b]) down to this rank's scenes.
    Returns (xyz_local, offset_local, *per_point_local, scene_ids)."""
    starts = torch.cat([offset.new_zeros(1), offset[:-1]]).tolist()
    ends = offset.tolist()
    ids = shard_scenes(len(ends), rank, world_size)
    sel = torch.cat([torch.arange(starts[i], ends[i]) for i in ids]) if ids else torch.zeros(0, dtype=torch.long)
    sel = sel.to(xyz.device)
    counts = torch.tensor([ends[i] - starts[i] for i in ids], dtype=offset.dtype, device=offset.device)
    return (xyz[sel], torch.cumsum(counts, 0).to(offset.dtype), *[t[sel] for t in per_point], ids)


def allreduce_gradients(grads: Sequence[torch.Tensor], group=None, average: bool = True, async_op: bool = False):
    """In-place sum (or mean) of the given gradient tensors over all ranks with ONE collective.

    async_op=True starts the collective and returns a `finish()` callable (None when there is nothing to reduce): the
    exchange of one layer's gradients then runs on the backend's own stream under the backward pass of the next layers,
    and `finish()` - called once at the end of the step - waits for it and writes the result back into the tensors."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None
    grads = [g for g in grads if g is not None]
    if not grads:
        return None
    flat = torch.cat([g.reshape(-1) for g in grads])
    work = dist.all_reduce(flat, group=group, async_op=async_op)

    def finish():
        if work is not None:
            work.wait()
        if average:
            flat.div_(dist.get_world_size(group))
        o = 0
        for g in grads:
            g.copy_(flat[o:o + g.numel()].view_as(g))
            o += g.numel()

    if async_op:
        return finish
    finish()
    return None


def global_throughput(points_local: int, seconds_local: float, device=None, group=None) -> float:
    """points/s of the whole job: total points over all ranks divided by the slowest rank's time."""
    if not dist.is_available() or not dist.is_initialized():
        return points_local / seconds_local
    t = torch.tensor([float(points_local), seconds_local], dtype=torch.float64, device=device)
    pts, sec = t[0:1].clone(), t[1:2].clone()
    dist.all_reduce(pts, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(sec, op=dist.ReduceOp.MAX, group=group)
    return float(pts.item() / sec.item())
