"""Row-sharded multi-GPU plumbing (one process per GPU, torch.distributed).

The lin_reg path shards by rows: every rank builds the moments of its shard, ONE all-reduce sums the
(p+t+1)^2 float64 moments, every rank solves redundantly and predicts its own rows.  This module holds the only pieces
that are not kernels: the row partition and the collective.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is
plumbing; no arithmetic of the hot path lives here.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_rows(n: int, rank: int, world: int, align: int = 128) -> Tuple[int, int]:
    """Contiguous [begin, end) row range of `rank`; boundaries are multiples of `align` (frame blocks) except the last."""
    blocks = (n + align - 1) // align
    per, extra = divmod(blocks, world)
    b0 = rank * per + min(rank, extra)
    b1 = b0 + per + (1 if rank < extra else 0)
    return min(b0 * align, n), min(b1 * align, n)


def allreduce_moments(M: torch.Tensor) -> torch.Tensor:
    """Sum the partial moments over all ranks in place (float64, (p+t+1)^2 values: latency-bound, ~9 KB at p = 32)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(M, op=dist.ReduceOp.SUM)
    return M
