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

from polars_ds_extension_b200.parallel import allreduce_moments, shard_rows


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
