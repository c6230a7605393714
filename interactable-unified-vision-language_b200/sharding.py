"""Data-parallel sharding of the encoder forward: images are independent (no BatchNorm, ``GroupNorm(1, C)`` is per
sample — ``sam/modeling/image_encoder.py`` has no cross-sample reduction), so a batch is split by image across ranks
with NO collective on the data path.  The reference's launcher does the same with one process per GPU
(``/root/reference/run:8-12``, ``utils/distributed.py``); gathering embeddings is optional and off the hot path."""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of ``total`` images owned by ``rank``; sizes differ by at most one (ragged batches)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_batch(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    b, e = shard_range(x.shape[0], rank, world)
    return x[b:e]


def all_gather_embeddings(local: Dict[str, torch.Tensor], total: int, group=None) -> Dict[str, torch.Tensor]:
    """OPTIONAL (off the hot path): collect every rank's embeddings, e.g. for evaluation.  Ragged shards are padded to the
    largest shard for the collective and trimmed afterwards."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]
    out: Dict[str, torch.Tensor] = {}
    for k, t in local.items():
        pad = torch.zeros((max(sizes),) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        bufs: List[torch.Tensor] = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad, group=group)
        out[k] = torch.cat([b[:n] for b, n in zip(bufs, sizes)], 0)
    return out
