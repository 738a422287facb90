"""N > 1 host logic on CPU: world_size-2 gloo run of the row-sharded protocol (shard -> partial moments -> one
all-reduce -> redundant solve).  The per-shard moments are built with numpy here (the CUDA kernels need a GPU); what is
under test is the partition and the collective the multi-GPU bench path relies on."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from polars_ds_extension_b200.parallel import (allreduce_moments, exclusive_prefix_moments, rolling_halo,
                                               shard_groups, shard_rows)


def test_shard_rows_partitions_exactly():
    for n in [1, 127, 128, 129, 1000, 100_000, 100_000_000]:
        for world in [1, 2, 3, 8]:
            spans = [shard_rows(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            assert all(s[0] % 128 == 0 for s in spans if s[0] < n)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, p, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(208)
    X = rng.standard_normal((n, p))
    y = X @ (((np.arange(p) % 7) - 3) / 4.0) + 0.1 * rng.standard_normal(n)
    b, e = shard_rows(n, rank, world)
    Z = np.column_stack([X[b:e], y[b:e], np.ones(e - b)])
    M = torch.from_numpy(Z.T @ Z)
    allreduce_moments(M)
    G = M[:p, :p].numpy()
    beta = np.linalg.solve(G, M[:p, p].numpy())
    if rank == 0:
        np.save(out, beta)
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process(tmp_path):
    n, p, world = 10_007, 6, 2
    out = str(tmp_path / "beta.npy")
    mp.spawn(_worker, args=(world, _free_port(), n, p, out), nprocs=world, join=True)
    rng = np.random.default_rng(208)
    X = rng.standard_normal((n, p))
    y = X @ (((np.arange(p) % 7) - 3) / 4.0) + 0.1 * rng.standard_normal(n)
    ref, *_ = np.linalg.lstsq(X, y, rcond=None)
    np.testing.assert_allclose(np.load(out), ref, rtol=1e-9, atol=1e-11)


def test_shard_groups_covers_every_group_once():
    rng = np.random.default_rng(3)
    for n_groups in [1, 2, 7, 1000]:
        sizes = rng.integers(1, 50, n_groups)
        off = np.concatenate([[0], np.cumsum(sizes)])
        for world in [1, 2, 3, 8]:
            spans = [shard_groups(off, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n_groups
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0] and a[0] <= a[1]
            if n_groups >= 100 * world:     # balanced by rows
                rows = [off[b] - off[a] for a, b in spans]
                assert max(rows) - min(rows) <= 2 * 50


def test_rolling_halo():
    assert rolling_halo(0, 1024) == 0
    assert rolling_halo(500, 1024) == 500
    assert rolling_halo(4096, 1024) == 1023
    assert rolling_halo(4096, 1) == 0


def _prefix_worker(rank, world, port, n, p, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    Z = np.column_stack([rng.standard_normal((n, p + 1)), np.ones(n)])
    b, e = shard_rows(n, rank, world)
    M = torch.from_numpy(Z[b:e].T @ Z[b:e])
    P = exclusive_prefix_moments(M)
    np.save(f"{out}_{rank}.npy", P.numpy())
    dist.destroy_process_group()


def test_two_rank_exclusive_prefix_moments(tmp_path):
    """recursive_lin_reg shard protocol: rank r receives the moments of all rows before its shard."""
    n, p, world = 5_003, 4, 2
    out = str(tmp_path / "prefix")
    mp.spawn(_prefix_worker, args=(world, _free_port(), n, p, out), nprocs=world, join=True)
    rng = np.random.default_rng(5)
    Z = np.column_stack([rng.standard_normal((n, p + 1)), np.ones(n)])
    for r in range(world):
        b, _ = shard_rows(n, r, world)
        np.testing.assert_allclose(np.load(f"{out}_{r}.npy"), Z[:b].T @ Z[:b], rtol=1e-12, atol=1e-9)
