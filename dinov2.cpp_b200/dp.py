"""Data-parallel host logic: one process per GPU, weights replicated, images sharded contiguously.

The reference has no multi-device path at all; images are independent, so the forward
pass needs no data-path collective.  The only exchange is the optional all-gather of the per-image
feature vectors ([cls] embeddings, or patch tokens) so that every rank ends up with the whole batch's
features — done with torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [start, stop) of rank `rank`: GPU g gets images [g*B, (g+1)*B) when world | n_items,
    otherwise the first n_items % world ranks take one extra image."""
    if world <= 0 or not (0 <= rank < world) or n_items < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_sizes(n_items: int, world: int) -> List[int]:
    return [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)]


def all_gather_features(local: torch.Tensor, n_items: int, group=None) -> torch.Tensor:
    """local: this rank's [n_local, ...] features (rows of shard_range).  Returns [n_items, ...] on every rank,
    in global image order.  Even shards use one all_gather_into_tensor; ragged shards are padded to the largest."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(n_items, world)
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} rows, expected {sizes[rank]}")
    local = local.contiguous()
    if len(set(sizes)) == 1:
        out = local.new_empty((n_items,) + tuple(local.shape[1:]))
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    m = max(sizes)
    padded = local.new_zeros((m,) + tuple(local.shape[1:]))
    padded[: local.shape[0]] = local
    buf = local.new_empty((world * m,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(buf, padded, group=group)
    return torch.cat([buf[r * m: r * m + sizes[r]] for r in range(world)], dim=0)
