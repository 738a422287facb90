"""Row-sharded rolling / recursive / grouped lin_reg on ONE GPU: the shards of a world-size-W run are
processed one after the other with exactly the calls each rank would make (halo rows for rolling, preceding-row
moments for recursive, group ranges for group_by) and must reproduce the single-shard result."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _frame(n, p, dtype, seed):
    import torch

    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    Z = torch.randn((p + 1, n), generator=g, device="cuda", dtype=dtype)
    Z[p] = Z[:p].sum(0) * 0.25 + 0.1 * Z[p]
    return Z


@pytest.mark.parametrize("dtype_name,n,p,bias,world", [("float32", 50_003, 4, True, 2), ("float64", 20_000, 3, False, 3),
                                                       ("float32", 9_000, 8, True, 4)])
def test_recursive_shards_continue_the_prefix(dtype_name, n, p, bias, world):
    import torch

    from polars_ds_extension_b200 import device as dev
    from polars_ds_extension_b200.parallel import shard_rows

    dtype = getattr(torch, dtype_name)
    Z = _frame(n, p, dtype, 11)
    Z[0, 777] = float("nan")          # a non-finite row never contributes, in any shard
    min_rows = p + int(bias) + 3
    full = dev.online_lin_reg(Z[:p], Z[p], window=0, min_rows=min_rows, add_bias=bias, l2_reg=0.01)
    prefix = torch.zeros((p + 2, p + 2), dtype=torch.float64, device="cuda")
    fin = torch.isfinite(Z).all(0)
    Zc = torch.where(fin[None, :], Z, torch.zeros_like(Z))
    for r in range(world):
        b, e = shard_rows(n, r, world)
        Xs, ys = Z[:p, b:e], Z[p, b:e]
        co, pr, va = dev.recursive_shard(Xs, ys, min_rows, prefix if r else None, b, add_bias=bias, l2_reg=0.01)
        assert torch.equal(va, full[2][b:e])
        ok = va.bool()
        tol = 2e-4 if dtype == torch.float32 else 1e-9
        assert torch.allclose(co[ok], full[0][b:e][ok], rtol=tol, atol=tol)
        assert torch.allclose(pr[ok], full[1][b:e][ok], rtol=tol, atol=tol, equal_nan=True)
        # what the next rank receives: moments of the finite rows seen so far (exclusive_prefix_moments in parallel.py)
        prefix = prefix + dev.moments(Zc[:p, b:e], Zc[p:p + 1, b:e], mask=fin[b:e].to(dtype).contiguous())
    assert int(full[2].sum()) == n - (min_rows - 1)


@pytest.mark.parametrize("dtype_name,n,p,window,world", [("float32", 40_001, 5, 300, 2), ("float64", 10_000, 2, 1024, 3)])
def test_rolling_shards_with_halo(dtype_name, n, p, window, world):
    import torch

    from polars_ds_extension_b200 import device as dev
    from polars_ds_extension_b200.parallel import rolling_halo, shard_rows

    dtype = getattr(torch, dtype_name)
    Z = _frame(n, p, dtype, 12)
    full = dev.online_lin_reg(Z[:p], Z[p], window=window, min_rows=window, add_bias=True)
    for r in range(world):
        b, e = shard_rows(n, r, world)
        h = rolling_halo(b, window)
        co, pr, va = dev.online_lin_reg(Z[:p, b - h:e], Z[p, b - h:e], window=window, min_rows=window, add_bias=True)
        assert torch.equal(va[h:], full[2][b:e])
        ok = va[h:].bool()
        tol = 2e-4 if dtype == torch.float32 else 1e-9
        assert torch.allclose(co[h:][ok], full[0][b:e][ok], rtol=tol, atol=tol)
        assert torch.allclose(pr[h:][ok], full[1][b:e][ok], rtol=tol, atol=tol)


def test_group_shards_match_single_launch():
    import torch

    from polars_ds_extension_b200 import device as dev
    from polars_ds_extension_b200.parallel import shard_groups

    rng = np.random.default_rng(4)
    sizes = rng.integers(20, 400, 500)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    n, p = int(off[-1]), 3
    Z = _frame(n, p, torch.float32, 13)
    offs = torch.from_numpy(off).cuda()
    full_beta, full_status = dev.grouped_lin_reg(Z[:p], Z[p], offs, add_bias=True)
    world = 3
    for r in range(world):
        g0, g1 = shard_groups(off, r, world)
        lo, hi = int(off[g0]), int(off[g1])
        local = torch.from_numpy(off[g0:g1 + 1] - off[g0]).cuda()
        beta, status = dev.grouped_lin_reg(Z[:p, lo:hi], Z[p, lo:hi], local, add_bias=True)
        assert torch.equal(status, full_status[g0:g1])
        assert torch.allclose(beta, full_beta[g0:g1], rtol=1e-6, atol=1e-7)
