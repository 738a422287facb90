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


def set_devices(devices) -> None:
    """Single process, several GPUs: every large lin_reg call of this process is row-sharded over `devices`
    (include/pdsb.h, pdsb_set_devices; the environment variable PDS_B200_DEVICES does the same at load time)."""
    import ctypes as C

    from ._lib import check, lib

    d = list(devices)
    arr = (C.c_int * max(len(d), 1))(*d)
    check(lib().pdsb_set_devices(arr, len(d)))


def init_world() -> int:
    """One process per GPU (torchrun): join the library's own NCCL communicator.  Rank 0 creates the 128-byte id,
    torch.distributed (whatever backend the launcher initialised) carries it to the other ranks, every rank joins.
    From here on plugin lin_reg calls are collective: one fit over the rows of all ranks.  Returns the world size."""
    import ctypes as C

    from ._lib import check, lib

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 1
    world, rank = dist.get_world_size(), dist.get_rank()
    buf = (C.c_uint8 * 128)()
    if rank == 0:
        check(lib().pdsb_comm_unique_id(buf))
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor(list(buf), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0)
    raw = bytes(t.cpu().tolist())
    check(lib().pdsb_comm_init_rank(world, rank, C.create_string_buffer(raw, 128)))
    return world


def destroy_world() -> None:
    from ._lib import lib

    lib().pdsb_comm_destroy()


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


def shard_groups(offsets, rank: int, world: int) -> Tuple[int, int]:
    """group_by: groups are independent, so ranks take contiguous group ranges balanced by ROW count; no collective on
    the data path.  `offsets` = the n_groups + 1 group boundaries of the key-sorted frame.  Returns the
    [g0, g1) group range of `rank`."""
    import numpy as np

    off = np.asarray(offsets, dtype=np.int64)
    n_groups = len(off) - 1
    if n_groups <= 0:
        return 0, 0
    total = int(off[-1] - off[0])
    # boundary r = first group whose start reaches r/world of the rows; monotone, covers every group exactly once
    cuts = [int(np.searchsorted(off[:-1] - off[0], (total * r) // world, side="left")) for r in range(world)] + [n_groups]
    cuts[0] = 0
    return cuts[rank], max(cuts[rank], cuts[rank + 1])


def rolling_halo(begin: int, window: int) -> int:
    """rolling_lin_reg on a row shard [begin, end): the shard also reads the `window - 1` rows before `begin`
    (a read-only overlap with its left neighbour, no collective); the outputs of the halo rows are dropped."""
    return min(begin, max(window - 1, 0))


def exclusive_prefix_moments(M: torch.Tensor) -> torch.Tensor:
    """recursive_lin_reg on row shards: every rank needs the moments of all rows before its shard.  One all-gather of the
    per-shard totals ((p+2)^2 float64 each), then a local sum over the lower ranks (deterministic order)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return torch.zeros_like(M)
    world, rank = dist.get_world_size(), dist.get_rank()
    parts = [torch.empty_like(M) for _ in range(world)]
    dist.all_gather(parts, M.contiguous())
    out = torch.zeros_like(M)
    for r in range(rank):
        out += parts[r]
    return out
