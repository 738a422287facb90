"""Parity tests proper: the CUDA path, called through the plugin C ABI exactly like Polars would call it
(pickled kwargs + Arrow C data -> `_polars_plugin_pl_lr*` in _polars_ds_b200.so), checked against

  (a) the reference's own known-answer tests (tests/cases.py: same seeds, external checkers, tolerances), and
  (b) the CPU oracle on the same seeded inputs (oracle/lin_reg_oracle.py), element by element.

Tolerances (north_star): 1e-6 relative for f64, 1e-4 relative for f32.
"""
import numpy as np
import pytest

import polars_ds_extension_b200 as pds
import polars_ds_extension_b200.config as cfg
from polars_ds_extension_b200 import Frame
from tests import cases
from tests.backends import OracleBackend, PluginBackend

pytestmark = pytest.mark.gpu

GPU = PluginBackend()
ORC = OracleBackend()


@pytest.mark.parametrize("case", cases.ALL_CASES, ids=lambda f: f.__name__)
def test_reference_cases_f64(case, monkeypatch):
    monkeypatch.setattr(cfg, "LIN_REG_EXPR_F64", True)
    case(GPU)


@pytest.mark.parametrize("case", cases.ALL_CASES, ids=lambda f: f.__name__)
def test_reference_cases_f32(case, monkeypatch):
    monkeypatch.setattr(cfg, "LIN_REG_EXPR_F64", False)
    case(GPU)


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    # vector-relative: max |a-b| over max |b| (a coefficient whose true value is ~0 has no meaningful own scale)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def _frame(seed, n, p, dtype=np.float64, noise=0.1):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, p))
    beta = ((np.arange(p) % 7) - 3) / 4.0
    y = X @ beta + noise * rng.standard_normal(n) + 0.5
    d = {f"x{i}": X[:, i].astype(dtype) for i in range(p)}
    d["y"] = y.astype(dtype)
    return Frame(d), [f"x{i}" for i in range(p)]


@pytest.mark.parametrize("f64", [True, False])
@pytest.mark.parametrize("n,p,bias", [(100_000, 4, True), (20_000, 32, False), (3_000, 64, True), (257, 1, True)])
def test_lin_reg_vs_oracle(monkeypatch, f64, n, p, bias):
    """config[0] (100k x 4 f64, bias) and scaled-down config[1] / config[4] shapes; coefficients and predictions."""
    monkeypatch.setattr(cfg, "LIN_REG_EXPR_F64", f64)
    df, xs = _frame(20 + p, n, p, np.float64 if f64 else np.float32)
    tol = 1e-6 if f64 else 1e-4
    e = pds.lin_reg(*xs, target="y", add_bias=bias)
    g, o = GPU.eval(df, e), ORC.eval(df, e)
    assert _rel(g, o) < tol
    # f32: the GPU (3xTF32 + f64 reduction) must be no farther from the f64 truth than the f32 oracle is
    if not f64:
        monkeypatch.setattr(cfg, "LIN_REG_EXPR_F64", True)
        truth = ORC.eval(df, pds.lin_reg(*xs, target="y", add_bias=bias))
        assert _rel(g, truth) <= max(2.0 * _rel(o, truth), 2e-6)
        monkeypatch.setattr(cfg, "LIN_REG_EXPR_F64", False)
    e = pds.lin_reg(*xs, target="y", add_bias=bias, return_pred=True)
    g, o = GPU.eval(df, e), ORC.eval(df, e)
    scale = np.abs(o["pred"][0]).max()
    assert np.max(np.abs(g["pred"][0] - o["pred"][0])) < tol * scale * 4
    assert np.max(np.abs(g["resid"][0] - o["resid"][0])) < tol * scale * 4


@pytest.mark.parametrize("f64", [True, False])
def test_solvers_and_ridge_vs_oracle(monkeypatch, f64):
    monkeypatch.setattr(cfg, "LIN_REG_EXPR_F64", f64)
    df, xs = _frame(31, 5000, 8, np.float64 if f64 else np.float32)
    tol = 1e-6 if f64 else 1e-4
    for solver in ["qr", "svd", "choleskey", "cholesky"]:
        for l2 in [0.0, 0.5]:
            e = pds.lin_reg(*xs, target="y", add_bias=True, solver=solver, l2_reg=l2)
            assert _rel(GPU.eval(df, e), ORC.eval(df, e)) < tol, (solver, l2)


@pytest.mark.parametrize("f64", [True, False])
def test_report_vs_oracle(monkeypatch, f64):
    monkeypatch.setattr(cfg, "LIN_REG_EXPR_F64", f64)
    df, xs = _frame(32, 4000, 5, np.float64 if f64 else np.float32)
    tol = 1e-6 if f64 else 2e-3
    for se in ["se", "hc0", "hc1", "hc2", "hc3"]:
        e = pds.lin_reg_report(*xs, target="y", add_bias=True, std_err=se)
        g, o = GPU.eval(df, e), ORC.eval(df, e)
        assert g["features"] == o["features"]
        for k in o:
            if k == "features":
                continue
            if k == "p>|t|":
                np.testing.assert_allclose(g[k], o[k], rtol=max(tol, 1e-5), atol=1e-30)
            else:
                np.testing.assert_allclose(g[k], o[k], rtol=tol, atol=tol * 1e-2, err_msg=f"{se}:{k}")


@pytest.mark.parametrize("f64", [True, False])
@pytest.mark.parametrize("window,p,bias,l2", [(1024, 8, False, 0.0), (37, 3, True, 0.1), (2, 1, False, 0.0), (5000, 4, True, 0.0)])
def test_rolling_vs_definition(monkeypatch, f64, window, p, bias, l2):
    """config[3] shape scaled down (window 1024, 8 features): every row == OLS on its window (the identity the
    reference's tests assert, test_linear_exprs.py:814-854)."""
    monkeypatch.setattr(cfg, "LIN_REG_EXPR_F64", f64)
    n = max(3 * window + 77, 4000)
    df, xs = _frame(40 + p, n, p, np.float64 if f64 else np.float32)
    r = GPU.eval(df, pds.rolling_lin_reg(*xs, target="y", window_size=window, add_bias=bias, l2_reg=l2))
    X = np.column_stack([df[c].to_numpy().astype(np.float64) for c in xs] + ([np.ones(n)] if bias else []))
    y = df["y"].to_numpy().astype(np.float64)
    assert all(c is None for c in r["coeffs"][: window - 1])
    from oracle.lin_reg_oracle import window_ols

    tol = 1e-6 if f64 else (2e-3 if window < 16 else 2e-4)
    rng = np.random.default_rng(1)
    rows = set(rng.integers(window - 1, n, 60).tolist()) | {window - 1, n - 1, window, min(n - 1, 2 * window)}
    for j in sorted(rows):
        ref = window_ols(X, y, j - window + 1, j + 1, lam=l2, add_bias=bias)
        assert _rel(r["coeffs"][j], ref) < tol, j
        assert abs(r["pred"][0][j] - X[j] @ ref) < tol * max(1.0, abs(X[j] @ ref)) * 10
    assert r["pred"][1][window - 1:].all() and not r["pred"][1][: window - 1].any()


@pytest.mark.parametrize("f64", [True, False])
def test_recursive_vs_definition(monkeypatch, f64):
    monkeypatch.setattr(cfg, "LIN_REG_EXPR_F64", f64)
    n, p = 6000, 4
    df, xs = _frame(50, n, p, np.float64 if f64 else np.float32)
    r = GPU.eval(df, pds.recursive_lin_reg(*xs, target="y", start_with=10, add_bias=True, l2_reg=0.01))
    X = np.column_stack([df[c].to_numpy().astype(np.float64) for c in xs] + [np.ones(n)])
    y = df["y"].to_numpy().astype(np.float64)
    from oracle.lin_reg_oracle import window_ols

    tol = 1e-6 if f64 else 2e-4
    for j in [9, 10, 50, 1023, 1024, 1025, 3000, n - 1]:
        ref = window_ols(X, y, 0, j + 1, lam=0.01, add_bias=True)
        assert _rel(r["coeffs"][j], ref) < (tol if j > 30 else tol * 50), j
    assert r["coeffs"][8] is None
    # and against the oracle's sequential Woodbury restatement on a short prefix (f64 only: f32 Woodbury drifts)
    if f64:
        o = ORC.eval(df.limit(300), pds.recursive_lin_reg(*xs, target="y", start_with=10, add_bias=True, l2_reg=0.01))
        for j in range(9, 300, 17):
            assert _rel(r["coeffs"][j], o["coeffs"][j]) < 1e-6


@pytest.mark.parametrize("f64", [True, False])
def test_grouped_config_shape(monkeypatch, f64):
    """config[2] scaled down: ragged contiguous groups x 8 features through the batched symbol == per-group oracle."""
    monkeypatch.setattr(cfg, "LIN_REG_EXPR_F64", f64)
    rng = np.random.default_rng(60)
    sizes = rng.integers(800, 1200, 40).tolist() + [9000, 20000, 9]
    gid = np.repeat(np.arange(len(sizes)), sizes)
    n = len(gid)
    df, xs = _frame(61, n, 8, np.float64 if f64 else np.float32)
    df = df.with_columns(g=gid)
    e = pds.lin_reg(*xs, target="y", add_bias=True)
    fast = GPU.group_eval(df, "g", e, fast=True)
    tol = 1e-6 if f64 else 2e-4
    for g in [0, 1, 17, 40, 41, 42]:
        ref = ORC.eval(df.filter(gid == g), e)
        assert _rel(fast[g], ref) < tol, g


def test_bit_reproducible(monkeypatch):
    monkeypatch.setattr(cfg, "LIN_REG_EXPR_F64", False)
    df, xs = _frame(70, 50_000, 16, np.float32)
    e = pds.lin_reg(*xs, target="y", add_bias=True, return_pred=True)
    a, b = GPU.eval(df, e), GPU.eval(df, e)
    assert np.array_equal(a["pred"][0], b["pred"][0])
