"""Multi-GPU layout of the path: independent clips, contiguous shards, one collective at the end.

Every clip (batch row) is independent through conditioning, sampling and decode, so rank r of
W takes rows [lo, hi) of the global batch with weights replicated, and the only exchange is an all-gather of the
330-d poses (B*128*330*4 bytes, 169 KB per clip) -- NCCL over NVLink on the GPUs, gloo in the CPU tests.
The reference itself only has nn.DataParallel scatter of the batch (train.py:94).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int):
    """Contiguous, balanced rows [lo, hi) of rank `rank`; the first (global_batch % world) ranks get one more."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, extra = divmod(global_batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_range(t.shape[0], rank, world)
    return t[lo:hi]


def gather_rows(local: torch.Tensor, global_batch: int, group=None) -> torch.Tensor:
    """All-gather per-rank row blocks (possibly ragged by one row) back into [global_batch, ...] in rank order."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    base, extra = divmod(global_batch, world)
    cap = base + (1 if extra else 0)
    lo, hi = shard_range(global_batch, rank, world)
    assert local.shape[0] == hi - lo, "local block does not match this rank's shard"
    padded = local
    if hi - lo < cap:
        padded = torch.cat([local, local.new_zeros((cap - (hi - lo),) + tuple(local.shape[1:]))])
    out = local.new_empty((world * cap,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    if not extra:
        return out
    parts = []
    for r in range(world):
        l, h = shard_range(global_batch, r, world)
        parts.append(out[r * cap: r * cap + (h - l)])
    return torch.cat(parts)
