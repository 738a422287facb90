"""Frozen golden vectors (tests/golden/lin_reg_golden.npz, written by tests/golden/make_golden.py from scikit-learn /
numpy / scipy — the external checkers of the reference's own tests) against BOTH evaluators of the same expressions:
the CPU oracle (`-m "not gpu"`) and the CUDA path through the plugin C ABI (`-m gpu`)."""
from pathlib import Path

import numpy as np
import pytest

import polars_ds_extension_b200 as pds
from polars_ds_extension_b200 import Frame

G = np.load(Path(__file__).parent / "golden" / "lin_reg_golden.npz")
FEATS = ("x1", "x2", "x3")


def _frame(X=None, **extra):
    X = G["X"] if X is None else X
    d = {"x1": X[:, 0], "x2": X[:, 1], "x3": X[:, 2], "y": G["y"], "y2": G["y2"], "w": G["w"], "k": G["keys"]}
    d.update(extra)
    return Frame(d)


def _check(be, f32, errs=None):
    """`errs` (optional list) receives the vector-relative error of every compared quantity, in call order."""
    import polars_ds_extension_b200.config as cfg

    assert cfg.LIN_REG_EXPR_F64 == (not f32)
    tol = 2e-4 if f32 else 1e-8
    df = _frame()

    def close(a, b, t=tol):
        a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
        if errs is not None:
            errs.append(float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)))
        np.testing.assert_allclose(a, b, rtol=t, atol=t)
    close(be.eval(df, pds.lin_reg(*FEATS, target="y", add_bias=True)), G["ols_bias"])
    close(be.eval(df, pds.lin_reg(*FEATS, target="y")), G["ols"])
    close(be.eval(df, pds.lin_reg(*FEATS, target="y", add_bias=True, l2_reg=0.1)), G["ridge_bias"], max(tol, 1e-7))
    for solver in ("svd", "choleskey", "qr"):
        close(be.eval(df, pds.lin_reg(*FEATS, target="y", add_bias=True, solver=solver)), G["ols_bias"])
    close(be.eval(df, pds.simple_lin_reg("x1", "y")), [G["X"][:, 0] @ G["y"] / (G["X"][:, 0] @ G["X"][:, 0])])
    # coordinate descent stops on max |delta beta| < tol: compare at the optimiser's own resolution
    close(be.eval(df, pds.lin_reg(*FEATS, target="y", add_bias=True, l1_reg=0.01, tol=1e-9, max_iter=20000)),
          G["lasso_bias"], 1e-3 if f32 else 1e-5)
    close(be.eval(df, pds.lin_reg(*FEATS, target="y", l1_reg=0.01, l2_reg=0.02, tol=1e-9, max_iter=20000)),
          G["enet"], 1e-3 if f32 else 1e-5)
    close(be.eval(_frame(G["Xn"]), pds.lin_reg(*FEATS, target="y", positive=True, tol=1e-10, max_iter=20000)),
          G["nnls"], 1e-3 if f32 else 1e-5)
    close(be.eval(df, pds.lin_reg(*FEATS, target="y", add_bias=True, weights="w")), G["wls_bias"])
    r = be.eval(df, pds.lin_reg_w_rcond(*FEATS, target="y"))
    close(r["coeffs"], G["rcond_coeffs"])
    close(r["singular_values"], G["rcond_sv"], max(tol, 1e-7))
    m = be.eval(df, pds.lin_reg(*FEATS, target=["y", "y2"], add_bias=True))
    close(np.vstack([m[k] for k in sorted(m)]), G["multi"])
    p = be.eval(df, pds.lin_reg(*FEATS, target="y", add_bias=True, return_pred=True))
    A = np.column_stack([G["X"], np.ones(len(G["y"]))])
    close(p["pred"][0], A @ G["ols_bias"], max(tol, 1e-7))
    close(p["resid"][0], G["y"] - A @ G["ols_bias"], max(tol, 1e-7))
    rep = be.eval(df, pds.lin_reg_report(*FEATS, target="y", add_bias=True))
    for row, key in enumerate(["beta", "std_err", "t", "p>|t|", "0.025", "0.975"]):
        np.testing.assert_allclose(rep[key], G["report"][row], rtol=5e-3 if f32 else 1e-6, atol=1e-30 if key == "p>|t|" else (1e-4 if f32 else 1e-10))
    head = df.slice(0, 400)
    roll = be.eval(head, pds.rolling_lin_reg(*FEATS, target="y", window_size=25))
    rec = be.eval(head, pds.recursive_lin_reg(*FEATS, target="y", start_with=10))
    for got, want, first in ((roll, G["rolling_w25"], 24), (rec, G["recursive_s10"], 9)):
        assert all(c is None for c in got["coeffs"][:first])
        # f32: a 25-row window of U(0,1) features has a condition number of ~1e2-1e3 on the Gram
        close(np.vstack(got["coeffs"][first:]), want[first:], 5e-2 if f32 else 1e-6)
    g = be.group_eval(df, "k", pds.lin_reg(*FEATS, target="y", add_bias=True))
    close(np.vstack(g), G["grouped_bias"], 5e-3 if f32 else 1e-7)       # the 7-row group is nearly square


@pytest.mark.parametrize("f32", [False, True], ids=["f64", "f32"])
def test_oracle_matches_golden(f32, monkeypatch):
    import polars_ds_extension_b200.config as cfg
    from tests.backends import OracleBackend

    monkeypatch.setattr(cfg, "LIN_REG_EXPR_F64", not f32)
    _check(OracleBackend(), f32)


@pytest.mark.gpu
@pytest.mark.parametrize("f32", [False, True], ids=["f64", "f32"])
def test_cuda_matches_golden(f32, monkeypatch):
    import polars_ds_extension_b200.config as cfg
    from tests.backends import PluginBackend

    monkeypatch.setattr(cfg, "LIN_REG_EXPR_F64", not f32)
    eg, eo = [], []
    _check(PluginBackend(), f32, eg)
    if f32:
        # SURVEY.md §8c, second half of the f32 rule: against the frozen (sklearn / numpy / scipy) answers the CUDA path
        # must be no farther off than the f32 oracle is — whatever absolute tolerance the shared checks above allow
        from tests.backends import OracleBackend

        _check(OracleBackend(), f32, eo)
        assert len(eg) == len(eo)
        worse = [(i, g, o) for i, (g, o) in enumerate(zip(eg, eo)) if g > max(o, 2e-6)]
        assert not worse, worse
    be = PluginBackend()
    fast = be.group_eval(_frame(), "k", pds.lin_reg(*FEATS, target="y", add_bias=True), fast=True)
    np.testing.assert_allclose(np.vstack(fast), G["grouped_bias"], rtol=5e-3 if f32 else 1e-7, atol=5e-3 if f32 else 1e-7)
