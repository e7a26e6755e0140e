"""Multi-GPU layout of the label path: collocation points are independent units, so each rank labels
a contiguous block of the global point index range and NO collective runs during label generation.  The only
exchange is the optional once-per-Picard-iteration all-gather of the finished (tx, y) label set for a trainer that
wants the reference's global-batch semantics."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [start, start+count) of the global point indices owned by `rank` (sizes differ by <= 1)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    q, r = divmod(int(n_total), int(world_size))
    start = rank * q + min(rank, r)
    return start, q + (1 if rank < r else 0)


def all_gather_labels(tx: torch.Tensor, y: torch.Tensor, n_total: int, group: Optional[dist.ProcessGroup] = None):
    """Assemble the global label set [n_total, 1+nx] x 2 on every rank from the rank-local blocks (rank order =
    point-index order).  NCCL on GPUs (NVLink/NVSwitch), gloo in the CPU tests."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    start, count = shard_range(n_total, rank, world)
    assert tx.shape[0] == count and y.shape[0] == count, (tx.shape, y.shape, count)
    cmax = shard_range(n_total, 0, world)[1]
    pad = torch.zeros(cmax, tx.shape[1] + y.shape[1], dtype=tx.dtype, device=tx.device)
    pad[:count, : tx.shape[1]] = tx
    pad[:count, tx.shape[1]:] = y
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    rows = [out[r][: shard_range(n_total, r, world)[1]] for r in range(world)]
    full = torch.cat(rows, 0)
    return full[:, : tx.shape[1]].contiguous(), full[:, tx.shape[1]:].contiguous()
