"""The reference's own lin_reg tests, lifted.

Every function restates one test of /root/reference/tests/test_linear_exprs.py or tests/test_many.py (cited), with the
same seed / literal data, the same external checker (scikit-learn, numpy.linalg, closed-form identities) and the same
tolerance.  Where the reference draws data from its unseeded ``pds.random`` plugin (out of scope here) the data is
re-created with ``numpy.random.default_rng`` and the same generating formula.  statsmodels is not in this image, so
the HC0-HC3 sandwich the reference checks against statsmodels is restated in numpy (``_hc_se``).

Each case takes a backend (tests/backends.py): the oracle (pinning the oracle to the reference's known answers) or the
CUDA plugin (the parity test proper).
"""
from __future__ import annotations

import numpy as np
import pytest

import polars_ds_extension_b200 as pds
import polars_ds_extension_b200.config as cfg
from polars_ds_extension_b200 import Frame


def _uniform_frame(seed, n, coefs, noise=1e-4, bias=0.0):
    rng = np.random.default_rng(seed)
    X = rng.random((n, len(coefs)))
    y = X @ np.asarray(coefs) + bias + rng.random(n) * noise
    d = {f"x{i + 1}": X[:, i] for i in range(X.shape[1])}
    d["y"] = y
    return Frame(d), X, y


def _f32():
    return not cfg.LIN_REG_EXPR_F64


# ---------------------------------------------------------------------------------------------------------
def case_lin_reg_against_sklearn(be):
    """test_lin_reg_against_sklearn (test_linear_exprs.py:61-120) / test_f32_lin_reg_against_sklearn (:313-373)."""
    from sklearn import linear_model

    df, x, y = _uniform_frame(101, 5000, [0.5, 0.1, -0.15])
    tol = 1e-4 if _f32() else 1e-5
    reg = linear_model.LinearRegression(fit_intercept=True).fit(x, y)
    c = be.eval(df, pds.lin_reg("x1", "x2", "x3", target="y", add_bias=True))
    assert np.all(np.abs(c[:3] - reg.coef_) < tol)
    assert abs(c[-1] - reg.intercept_) < max(tol, 1e-5 * abs(reg.intercept_) + 1e-8)
    reg = linear_model.Ridge(alpha=0.1, fit_intercept=True).fit(x, y)
    c = be.eval(df, pds.lin_reg("x1", "x2", "x3", target="y", l2_reg=0.1, add_bias=True))
    assert np.all(np.abs(c[:3] - reg.coef_) < 1e-3)
    assert abs(c[-1] - reg.intercept_) < 1e-3
    if _f32():
        reg = linear_model.LinearRegression(fit_intercept=True).fit(x, y)
        p = be.eval(df, pds.lin_reg("x1", "x2", "x3", target="y", add_bias=True, return_pred=True))
        assert np.all(np.abs(p["pred"][0] - reg.predict(x)) < 1e-3)
        rep = be.eval(df, pds.lin_reg_report("x1", "x2", "x3", target="y", add_bias=True))
        assert np.all(np.isfinite(rep["beta"]))
        assert np.all(np.abs(rep["beta"][:3] - reg.coef_) < 1e-3)


def _rolling_case(be, l2):
    """test_rolling_ridge (:123-166) / test_rolling_lin_reg (:814-854): rolling == per-window lin_reg."""
    df, _, _ = _uniform_frame(102, 500, [0.5, 0.25, -0.15])
    rtol = 2e-3 if _f32() else 1e-5        # reference uses assert_frame_equal defaults (rel 1e-5) in f64
    for w in [5, 8, 12, 15]:
        res = be.eval(df, pds.rolling_lin_reg("x1", "x2", "x3", target="y", l2_reg=l2, window_size=w))
        assert all(c is None for c in res["coeffs"][: w - 1])
        for i in range(len(df) - w + 1):
            ans = be.eval(df.slice(i, w), pds.lin_reg("x1", "x2", "x3", l2_reg=l2, target="y"))
            got = res["coeffs"][i + w - 1]
            np.testing.assert_allclose(got, ans, rtol=rtol, atol=1e-8 if not _f32() else 2e-3)


def case_rolling_ridge(be):
    _rolling_case(be, 0.1)


def case_rolling_lin_reg(be):
    _rolling_case(be, 0.0)


def _hc_se(X, y, kind):
    """numpy restatement of statsmodels OLS(...).bse / HC{0..3}_se used by test_hc_lin_reg_report (:169-201)."""
    n, k = X.shape
    G = np.linalg.inv(X.T @ X)
    beta = G @ X.T @ y
    e = y - X @ beta
    if kind == "se":
        return np.sqrt(np.diag(G) * (e @ e) / (n - k))
    h = np.einsum("ij,jk,ik->i", X, G, X)
    w = {"hc0": e**2, "hc1": e**2 * n / (n - k), "hc2": e**2 / (1 - h), "hc3": e**2 / (1 - h) ** 2}[kind]
    return np.sqrt(np.diag(G @ (X.T * w) @ X @ G))


def case_hc_lin_reg_report(be):
    """test_hc_lin_reg_report (:169-201); tolerance 1e-7 (the reference's one-sided check made two-sided)."""
    rng = np.random.default_rng(103)
    X = rng.random((1000, 3))
    y = X[:, 0] * 0.15 + X[:, 1] * 0.3 + 0.1 + rng.standard_normal(1000) * 0.05
    df = Frame({"x1": X[:, 0], "x2": X[:, 1], "x3": X[:, 2], "target": y})
    for se_type in ("se", "hc0", "hc1", "hc2", "hc3"):
        r = be.eval(df, pds.lin_reg_report("x1", "x2", "x3", target="target", std_err=se_type))
        key = "std_err" if se_type == "se" else f"{se_type}_se"
        tol = 1e-4 if _f32() else 1e-7
        assert np.all(np.abs(r[key] - _hc_se(X, y, se_type)) < tol), se_type
        assert r["features"] == ["x1", "x2", "x3"]


def case_report_values(be):
    """Full report row check (beta, t, p, CI, r2, adj_r2) against scipy/numpy closed forms; covers a16/a17."""
    from scipy import stats

    rng = np.random.default_rng(2)
    n = 300
    x1, x2 = rng.standard_normal(n), rng.standard_normal(n)
    y = 0.5 * x1 - 0.3 * x2 + 0.1 * rng.standard_normal(n)
    df = Frame({"x1": x1, "x2": x2, "y": y})
    r = be.eval(df, pds.lin_reg_report("x1", "x2", target="y", add_bias=True))
    X = np.column_stack([x1, x2, np.ones(n)])
    beta, *_ = np.linalg.lstsq(X, y, rcond=None)
    tol = 1e-4 if _f32() else 1e-9
    np.testing.assert_allclose(r["beta"], beta, rtol=tol, atol=tol)     # :984-1028 (rtol 1e-10 there)
    e = y - X @ beta
    dof = n - 3
    se = np.sqrt(np.diag(np.linalg.inv(X.T @ X)) * (e @ e) / dof)
    np.testing.assert_allclose(r["std_err"], se, rtol=max(tol, 1e-8))
    t = beta / se
    np.testing.assert_allclose(r["t"], t, rtol=1e-3 if _f32() else 1e-7)
    np.testing.assert_allclose(r["p>|t|"], 2 * stats.t.sf(np.abs(t), dof), rtol=2e-3 if _f32() else 1e-6, atol=1e-30 if _f32() else 1e-300)
    ta = stats.t.ppf(0.975, dof)
    np.testing.assert_allclose(r["0.025"], beta - ta * se, rtol=1e-3 if _f32() else 1e-7, atol=1e-6 if _f32() else 1e-12)
    np.testing.assert_allclose(r["0.975"], beta + ta * se, rtol=1e-3 if _f32() else 1e-7, atol=1e-6 if _f32() else 1e-12)
    ratio = (e @ e) / (np.var(y, ddof=1) * n)                            # reference quirk kept (:867)
    assert abs(r["r2"] - (1 - ratio)) < (1e-5 if _f32() else 1e-10)
    assert abs(r["adj_r2"] - (1 - ratio * (n - 1) / (dof - 1))) < (1e-5 if _f32() else 1e-10)
    assert r["features"] == ["x1", "x2", "__bias__"]


def case_f32_everything_runs(be):
    """test_f32_lin_reg (:204-310): every f32 symbol runs (and here: returns finite, right-shaped output)."""
    old = cfg.LIN_REG_EXPR_F64
    cfg.LIN_REG_EXPR_F64 = False
    try:
        rng = np.random.default_rng(104)
        n = 500
        X = rng.random((n, 3))
        y = X @ [0.5, 0.25, -0.15] + rng.random(n) * 1e-4
        y2 = X @ [1.0, 0.3, -0.1]
        df = Frame({"x1": X[:, 0], "x2": X[:, 1], "x3": X[:, 2], "y": y, "y2": y2})
        xs = ("x1", "x2", "x3")
        assert be.eval(df, pds.lin_reg(*xs, target="y")).shape == (3,)
        m = be.eval(df, pds.lin_reg(*xs, target=["y", "y2"]))
        assert list(m) == ["target_0", "target_1"] and m["target_1"].shape == (3,)
        assert be.eval(df, pds.lin_reg(*xs, target="y", return_pred=True))["pred"][0].shape == (n,)
        assert np.all(np.isfinite(be.eval(df, pds.lin_reg(*xs, target="y", l1_reg=0.01))))
        assert be.eval(df, pds.lin_reg(*xs, target="y", l2_reg=0.01, return_pred=True))["resid"][0].shape == (n,)
        assert np.all(np.isfinite(be.eval(df, pds.lin_reg(*xs, target="y", l1_reg=0.01, l2_reg=0.01))))
        assert be.eval(df, pds.lin_reg(*xs, target=["y", "y2"], l2_reg=0.01))["target_0"].shape == (3,)
        assert be.eval(df, pds.lin_reg(*xs, target=["y", "y2"], l2_reg=0.01, null_policy="0.1"))["target_0"].shape == (3,)
        assert np.all(np.isfinite(be.eval(df, pds.lin_reg_report(*xs, target="y"))["beta"]))
        assert np.all(np.isfinite(be.eval(df, pds.lin_reg_report(*xs, target="y", weights="x1"))["beta"]))
        r = be.eval(df, pds.lin_reg_w_rcond(*xs, target="y", rcond=0.3))
        assert r["coeffs"].shape == (3,) and r["singular_values"].shape == (3,)
        r = be.eval(df, pds.rolling_lin_reg(*xs, target="y", window_size=3))
        assert r["coeffs"][1] is None and r["coeffs"][2] is not None
        r = be.eval(df, pds.recursive_lin_reg(*xs, target="y", start_with=3))
        assert r["coeffs"][1] is None and r["coeffs"][2] is not None
    finally:
        cfg.LIN_REG_EXPR_F64 = old


def case_multi_pred_correctness(be):
    """test_pl_lr_multi_pred_correctness (:376-408), seed 42, atol 1e-8 (f64)."""
    rng = np.random.default_rng(42)
    n = 1000
    X = rng.standard_normal((n, 5))
    tb = np.array([[0.5, -0.2, 0.1, 0.3, -0.4], [0.1, 0.6, -0.3, 0.0, 0.2], [-0.2, 0.0, 0.7, -0.1, 0.3]])
    Y = X @ tb.T + 0.01 * rng.standard_normal((n, 3))
    df = Frame({f"x{i + 1}": X[:, i] for i in range(5)} | {"y1": Y[:, 0], "y2": Y[:, 1], "y3": Y[:, 2]})
    feats = [f"x{i + 1}" for i in range(5)]
    multi = be.eval(df, pds.lin_reg(*feats, target=["y1", "y2", "y3"], return_pred=True))
    for i, t in enumerate(["y1", "y2", "y3"]):
        single = be.eval(df, pds.lin_reg(*feats, target=t, return_pred=True))
        np.testing.assert_allclose(multi[f"target_{i}_pred"][0], single["pred"][0], rtol=0, atol=1e-4 if _f32() else 1e-8)


def case_skip_null_literal(be):
    """test_lin_reg_skip_null (:411-432): literal frame, null row -> null pred/resid, exact fit elsewhere."""
    df = Frame({"y": [None, 9.5, 10.5, 11.5, 12.5], "a": [1, 9, 10, 11, 12], "b": [1.0, 0.5, 0.5, 0.5, 0.5]})
    r = be.eval(df, pds.lin_reg("a", "b", target="y", return_pred=True, null_policy="skip"))
    assert list(r["pred"][1]) == [False, True, True, True, True]
    assert list(r["resid"][1]) == [False, True, True, True, True]
    tol = 1e-3 if _f32() else 1e-9
    np.testing.assert_allclose(r["pred"][0][1:], [9.5, 10.5, 11.5, 12.5], atol=tol)
    np.testing.assert_allclose(r["resid"][0][1:], 0.0, atol=tol)


def case_group_by_literal(be):
    """test_lin_reg_in_group_by (:435-474): group_by result == filtered result (rank-deficient y=1 fit, gate off for
    exactness of the comparison is not needed: both sides run the same path)."""
    df = Frame({"A": [1] * 4 + [2] * 4, "Y": [1.0] * 8, "X1": [1, 2, 3, 4, 5, 6, 7, 8], "X2": [2, 3, 4, 1, 6, 7, 8, 5]})
    e = pds.lin_reg("X1", "X2", target="Y", add_bias=False, return_pred=True)
    keys = np.array([1] * 4 + [2] * 4)
    grouped = be.group_eval(df, "A", e)
    for gi, k in enumerate([1, 2]):
        alone = be.eval(df.filter(keys == k), e)
        np.testing.assert_allclose(grouped[gi]["pred"][0], alone["pred"][0], rtol=1e-6)
        # and against the normal equations directly
        X = np.column_stack([np.array(df["X1"].to_pylist(), float)[keys == k], np.array(df["X2"].to_pylist(), float)[keys == k]])
        beta = np.linalg.solve(X.T @ X, X.T @ np.ones(4))
        np.testing.assert_allclose(alone["pred"][0], X @ beta, rtol=1e-4 if _f32() else 1e-9)


def case_rcond(be):
    """test_lin_reg_with_rcond (:477-512): vs np.linalg.lstsq(rcond=0.3), 1e-10 (f64)."""
    rng = np.random.default_rng(105)
    X = rng.random((5000, 3))
    y = X[:, 0] + X[:, 1] * 0.2 - 0.3 * X[:, 2]
    df = Frame({"x1": X[:, 0], "x2": X[:, 1], "x3": X[:, 2], "y": y})
    np_coeffs, _, _, np_svs = np.linalg.lstsq(X, y, rcond=0.3)
    r = be.eval(df, pds.lin_reg_w_rcond("x1", "x2", "x3", target="y", rcond=0.3))
    tol = 1e-3 if _f32() else 1e-10
    assert np.all(np.abs(r["coeffs"] - np_coeffs) < tol)
    assert np.all(np.abs(r["singular_values"] - np_svs) < (1e-2 if _f32() else 1e-10))


def case_rcond_truncates(be):
    """test_lin_reg_with_rcond_truncates_singular_value (:515-554), seed 123."""
    if _f32():
        pytest.skip("reference runs this case in f64 only")
    rng = np.random.default_rng(123)
    n = 2000
    x1 = rng.standard_normal(n)
    x2 = x1 + rng.standard_normal(n) * 1e-6
    x3 = rng.standard_normal(n)
    y = x1 + 0.5 * x2 - 0.3 * x3
    df = Frame({"x1": x1, "x2": x2, "x3": x3, "y": y})
    rcond = 1e-3
    r = be.eval(df, pds.lin_reg_w_rcond("x1", "x2", "x3", target="y", rcond=rcond))
    X = np.column_stack([x1, x2, x3])
    evals, evecs = np.linalg.eigh(X.T @ X)
    thr = rcond * np.sqrt(evals.max())
    assert (evals < thr).any()
    pinv = sum((1.0 / ev) * np.outer(v, v) for ev, v in zip(evals, evecs.T) if ev >= thr)
    w_ref = pinv @ (X.T @ y)
    assert np.all(np.isfinite(r["coeffs"]))
    np.testing.assert_allclose(r["coeffs"], w_ref, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(np.sort(r["singular_values"])[::-1], np.sqrt(np.sort(evals))[::-1], rtol=1e-6, atol=1e-6)


def case_lasso(be):
    """test_lasso_regression (:557-604): vs sklearn Lasso, 1e-4."""
    from sklearn import linear_model

    df, x, y = _uniform_frame(106, 5000, [0.5, 0.25, -0.15])
    for lam in [0.01, 0.05, 0.1, 0.2]:
        c = be.eval(df, pds.lin_reg("x1", "x2", "x3", target="y", l1_reg=lam, add_bias=False))
        sk = linear_model.Lasso(alpha=lam, fit_intercept=False).fit(x, y)
        assert np.all(np.abs(sk.coef_ - c) < 1e-4)
    for lam in [0.01, 0.05, 0.1, 0.2]:
        c = be.eval(df, pds.lin_reg("x1", "x2", "x3", target="y", l1_reg=lam, add_bias=True))
        sk = linear_model.Lasso(alpha=lam, fit_intercept=True).fit(x, y)
        assert np.all(np.abs(sk.coef_ - c[:3]) < 1e-4)
        assert abs(c[-1] - sk.intercept_) < 1e-4


def case_positive(be):
    """test_positive_lin_reg (:607-674): NNLS vs sklearn LinearRegression(positive=True) 1e-5; elastic net 1e-4."""
    from sklearn.linear_model import ElasticNet, LinearRegression

    df, x, y = _uniform_frame(107, 5000, [0.5, 0.25, -0.15])
    for bias in [True, False]:
        c = be.eval(df, pds.lin_reg("x1", "x2", "x3", target="y", positive=True, add_bias=bias))
        assert np.all((c if not bias else c[:-1]) >= 0.0)
        sk = LinearRegression(positive=True, fit_intercept=bias).fit(x, y)
        tol = 2e-4 if _f32() else 1e-5   # f32 twin stops after 200 sweeps (linear_regression_f32.rs:343)
        assert np.all(np.isclose(c[:3], sk.coef_, atol=tol))
        if bias:
            assert np.isclose(float(c[-1]), sk.intercept_, atol=tol)
    for reg, bias in zip([0.01, 0.05, 0.1, 0.2], [False, True, False, True]):
        c = be.eval(df, pds.lin_reg("x1", "x2", "x3", target="y", l1_reg=reg, l2_reg=reg, add_bias=bias))
        sk = ElasticNet(alpha=2 * reg, l1_ratio=0.5, fit_intercept=bias).fit(x, y)
        assert np.all(np.isclose(c[:3], sk.coef_, atol=1e-4))
        if bias:
            assert np.isclose(float(c[-1]), sk.intercept_, atol=1e-4)


def case_elastic_net(be):
    """test_elastic_net_regression (:677-715)."""
    from sklearn import linear_model

    df, x, y = _uniform_frame(108, 5000, [0.5, 0.25, -0.15])
    for reg in [0.01, 0.05, 0.1, 0.2]:
        c = be.eval(df, pds.lin_reg("x1", "x2", "x3", target="y", l1_reg=reg, l2_reg=reg, add_bias=False))
        sk = linear_model.ElasticNet(alpha=2 * reg, l1_ratio=0.5, fit_intercept=False).fit(x, y)
        assert np.all(np.abs(sk.coef_ - c) < 1e-4)


def _recursive_case(be, l2):
    """test_recursive_lin_reg / test_recursive_ridge (:718-811): recursive[i-1] == lin_reg(first i rows), 1e-5."""
    df, _, _ = _uniform_frame(109, 1000, [0.5, 0.25, -0.15])
    start = 3
    r = be.eval(df, pds.recursive_lin_reg("x1", "x2", "x3", target="y", l2_reg=l2, start_with=start))
    assert r["coeffs"][0] is None and r["coeffs"][1] is None
    # the first rows are a nearly exactly-determined 3x3 system: conditioning is what limits f32 here
    tol = 5e-2 if _f32() else 1e-5
    for i in range(start, 30):
        normal = be.eval(df.limit(i), pds.lin_reg("x1", "x2", "x3", target="y", l2_reg=l2, singular_x_tol=0.0))
        assert np.all(np.abs(normal - r["coeffs"][i - 1]) < tol), i


def case_recursive_lin_reg(be):
    _recursive_case(be, 0.0)


def case_recursive_ridge(be):
    _recursive_case(be, 0.1)


def case_rolling_null_skips(be):
    """test_rolling_null_skips (:858-908): null pattern of the skip-window output."""
    rng = np.random.default_rng(110)
    n = 1000
    X = rng.random((n, 3))
    nulls = rng.random((n, 3)) < 0.15
    y = X @ [0.15, 0.3, -1.5] + rng.random(n) * 1e-4
    cols = {}
    for j in range(3):
        cols[f"x{j + 1}"] = [None if nulls[i, j] else float(X[i, j]) for i in range(n)]
    anynull = nulls.any(axis=1)
    cols["y"] = [None if anynull[i] else float(y[i]) for i in range(n)]
    df = Frame(cols)
    w, mv = 6, 5
    r = be.eval(df, pds.rolling_lin_reg("x1", "x2", "x3", target="y", window_size=w, min_valid_rows=mv, null_policy="skip"))
    should = [True] * (w - 1)
    for i in range(n - w + 1):
        should.append((w - anynull[i:i + w].sum()) < mv)
    assert [c is None for c in r["coeffs"]] == should


def case_many_small_groups(be):
    """test_lin_reg_many_small_groups_matches_per_group (:918-953), seed 0: grouped == per-group, 1e-12 (f64)."""
    rng = np.random.default_rng(0)
    n_groups, n_per = 200, 25
    gids = np.repeat(np.arange(n_groups), n_per)
    df = Frame({"g": gids, "x1": rng.standard_normal(len(gids)), "x2": rng.standard_normal(len(gids)),
                "y": rng.standard_normal(len(gids))})
    e = pds.lin_reg("x1", "x2", target="y", add_bias=True)
    fast = be.group_eval(df, "g", e, fast=True)
    slow = be.group_eval(df.slice(0, 10 * n_per), "g", e, fast=False)
    tol = 1e-4 if _f32() else 1e-10
    for g in range(n_groups):
        sub = df.filter(gids == g)
        X = np.column_stack([sub["x1"].to_numpy(), sub["x2"].to_numpy(), np.ones(n_per)])
        ref, *_ = np.linalg.lstsq(X, sub["y"].to_numpy(), rcond=None)
        np.testing.assert_allclose(fast[g], ref, rtol=tol, atol=tol)
        if g < 10:
            np.testing.assert_allclose(slow[g], ref, rtol=tol, atol=tol)


def case_bias_equivalence(be):
    """test_lin_reg_with_bias_appended_column_equivalence (:956-981), seed 1."""
    rng = np.random.default_rng(1)
    n = 500
    df = Frame({"x1": rng.standard_normal(n), "x2": rng.standard_normal(n), "y": rng.standard_normal(n), "ones": np.ones(n)})
    a = be.eval(df, pds.lin_reg("x1", "x2", target="y", add_bias=True))
    b = be.eval(df, pds.lin_reg("x1", "x2", "ones", target="y", add_bias=False))
    np.testing.assert_allclose(a, b, rtol=1e-4 if _f32() else 1e-10, atol=1e-6 if _f32() else 1e-12)


def case_report_cast_guard(be):
    """test_lin_reg_report_already_float64_cast_guard (:984-1028), seed 2: f64 vs f32-typed inputs, vs lstsq."""
    rng = np.random.default_rng(2)
    n = 300
    x1, x2 = rng.standard_normal(n), rng.standard_normal(n)
    y = 0.5 * x1 - 0.3 * x2 + 0.1 * rng.standard_normal(n)
    d64 = Frame({"x1": x1, "x2": x2, "y": y})
    d32 = Frame({"x1": x1.astype(np.float32), "x2": x2.astype(np.float32), "y": y.astype(np.float32)})
    r64 = be.eval(d64, pds.lin_reg_report("x1", "x2", target="y", add_bias=True))
    r32 = be.eval(d32, pds.lin_reg_report("x1", "x2", target="y", add_bias=True))
    np.testing.assert_allclose(r64["beta"], r32["beta"], rtol=1e-3 if _f32() else 1e-6, atol=1e-4 if _f32() else 1e-7)
    X = np.column_stack([x1, x2, np.ones(n)])
    ref, *_ = np.linalg.lstsq(X, y, rcond=None)
    np.testing.assert_allclose(r64["beta"], ref, rtol=1e-4 if _f32() else 1e-10, atol=1e-5 if _f32() else 1e-12)


def case_wls_multichunk(be):
    """test_wls_report_multichunked_weights_dont_panic (:1031-1066), seed 3."""
    import pyarrow as pa

    rng = np.random.default_rng(3)
    n = 200
    x = rng.standard_normal(n)
    y = 2.0 * x + 0.1 * rng.standard_normal(n)
    w = rng.uniform(0.5, 1.5, n)
    w_multi = pa.chunked_array([pa.array(w[: n // 2]), pa.array(w[n // 2:])])
    assert w_multi.num_chunks == 2
    e = pds.lin_reg_report("x", target="y", weights="w", add_bias=True)
    r = be.eval(Frame({"x": x, "y": y, "w": w_multi}), e)
    r1 = be.eval(Frame({"x": x, "y": y, "w": w}), e)
    np.testing.assert_allclose(r["beta"], r1["beta"], rtol=1e-12, atol=1e-12)
    X = np.column_stack([x, np.ones(n)])
    ref = np.linalg.solve((X.T * w) @ X, (X.T * w) @ y)
    np.testing.assert_allclose(r["beta"], ref, rtol=1e-4 if _f32() else 1e-10)


def case_multi_target_struct(be):
    """test_lin_reg_multi_target_struct_output (:1069-1113), seed 4: field names + per-target equality 1e-12."""
    rng = np.random.default_rng(4)
    n = 1000
    df = Frame({"x1": rng.standard_normal(n), "x2": rng.standard_normal(n), "y1": rng.standard_normal(n), "y2": rng.standard_normal(n)})
    m = be.eval(df, pds.lin_reg("x1", "x2", target=["y1", "y2"], add_bias=True))
    assert list(m.keys()) == ["target_0", "target_1"]
    c1 = be.eval(df, pds.lin_reg("x1", "x2", target="y1", add_bias=True))
    c2 = be.eval(df, pds.lin_reg("x1", "x2", target="y2", add_bias=True))
    tol = 1e-5 if _f32() else 1e-12
    np.testing.assert_allclose(m["target_0"], c1, rtol=tol, atol=tol)
    np.testing.assert_allclose(m["target_1"], c2, rtol=tol, atol=tol)


def case_single_big_fit(be):
    """test_lin_reg_single_big_fit_no_regression_path (:1116-1142), seed 5: 50 000 x 6 + bias vs sklearn rtol 1e-8."""
    from sklearn.linear_model import LinearRegression

    rng = np.random.default_rng(5)
    n, p = 50_000, 6
    X = rng.standard_normal((n, p))
    y = X @ np.array([0.4, -0.2, 0.7, 0.0, -0.1, 0.3]) + 0.05 * rng.standard_normal(n)
    df = Frame({f"x{i}": X[:, i] for i in range(p)} | {"y": y})
    c = be.eval(df, pds.lin_reg(*[f"x{i}" for i in range(p)], target="y", add_bias=True))
    sk = LinearRegression(fit_intercept=True).fit(X, y)
    tol = 1e-4 if _f32() else 1e-8
    np.testing.assert_allclose(c, list(sk.coef_) + [sk.intercept_], rtol=tol, atol=1e-5 if _f32() else 1e-10)


def case_null_skip_small_groups(be):
    """test_lin_reg_null_skip_in_small_group (:1145-1176): literal frame, skip-null inside group_by."""
    df = Frame({"g": [1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3],
                "x": [1.0, 2.0, None, 4.0, 1.0, None, 3.0, 4.0, 1.0, 2.0, 3.0, 4.0],
                "y": [2.0, 4.0, 6.0, 8.0, 1.0, 2.0, 3.0, None, 0.5, 1.0, 1.5, 2.0]})
    e = pds.lin_reg("x", target="y", add_bias=True, null_policy="skip", singular_x_tol=0.0)
    out = be.group_eval(df, "g", e)
    fast = be.group_eval(df, "g", e, fast=True)
    keys = np.array(df["g"].to_pylist())
    for gi, g in enumerate([1, 2, 3]):
        sub = df.filter(keys == g).drop_nulls()
        X = np.column_stack([np.array(sub["x"].to_pylist()), np.ones(len(sub))])
        ref, *_ = np.linalg.lstsq(X, np.array(sub["y"].to_pylist()), rcond=None)
        tol = 1e-4 if _f32() else 1e-9
        np.testing.assert_allclose(out[gi], ref, rtol=tol, atol=tol)
        np.testing.assert_allclose(fast[gi], ref, rtol=tol, atol=tol)


# ---- singular_x_tol gate (:1184-1340) ------------------------------------------------------------------
def _collinear(n=64, seed=0):
    rng = np.random.default_rng(seed)
    x1 = rng.standard_normal(n)
    return Frame({"x1": x1, "x2": 2.0 * x1, "y": rng.standard_normal(n)})


def case_gate_collinear_nulls(be):
    assert be.eval(_collinear(), pds.lin_reg("x1", "x2", target="y", add_bias=False)) is None          # :1204-1208


def case_gate_off_finite(be):
    c = be.eval(_collinear(), pds.lin_reg("x1", "x2", target="y", add_bias=False, singular_x_tol=0.0))  # :1211-1219
    assert c is not None and len(c) == 2


def case_gate_well_conditioned(be):
    from sklearn.linear_model import LinearRegression                                                   # :1222-1246

    rng = np.random.default_rng(7)
    n = 500
    X = rng.standard_normal((n, 3))
    y = X[:, 0] - 0.5 * X[:, 1] + 2.0 * X[:, 2]
    df = Frame({"x1": X[:, 0], "x2": X[:, 1], "x3": X[:, 2], "y": y})
    c = be.eval(df, pds.lin_reg("x1", "x2", "x3", target="y", add_bias=False))
    assert c is not None
    ref = LinearRegression(fit_intercept=False).fit(X, y).coef_
    tol = 1e-4 if _f32() else 1e-9
    np.testing.assert_allclose(c, ref, rtol=tol, atol=tol)


def case_gate_group_by(be):
    rng = np.random.default_rng(11)                                                                     # :1249-1268
    n = 40
    good_x1, bad_x1 = rng.standard_normal(n), rng.standard_normal(n)
    df = Frame({"g": ["good"] * n + ["bad"] * n, "x1": np.concatenate([good_x1, bad_x1]),
                "x2": np.concatenate([rng.standard_normal(n), 2.0 * bad_x1]), "y": rng.standard_normal(2 * n)})
    e = pds.lin_reg("x1", "x2", target="y", add_bias=False)
    for fast in (False, True):
        res = be.group_eval(df, "g", e, fast=fast)
        assert res[0] is not None and res[1] is None


def case_gate_return_pred(be):
    df = _collinear(n=32)                                                                               # :1271-1278
    r = be.eval(df, pds.lin_reg("x1", "x2", target="y", add_bias=False, return_pred=True))
    assert (~r["pred"][1]).sum() == 32 and (~r["resid"][1]).sum() == 32


def case_gate_multi_target(be):
    df = _collinear(n=64)                                                                               # :1281-1288
    y = np.array(df["y"].to_pylist())
    df = df.with_columns(y2=y * 0.5 + 1.0)
    s = be.eval(df, pds.lin_reg("x1", "x2", target=["y", "y2"], add_bias=False))
    assert s["target_0"] is None and s["target_1"] is None


def _scaled_singular(n=2000, feats=7, scale=1e3, seed=3):
    rng = np.random.default_rng(seed)
    base = rng.standard_normal(n) * scale
    cols = {f"x{i}": base * (i + 1) for i in range(feats)}
    cols["y"] = rng.standard_normal(n) * scale
    return Frame(cols), [f"x{i}" for i in range(feats)]


def _scaled_full_rank(n=2000, feats=7, scale=1e3, seed=4):
    rng = np.random.default_rng(seed)
    cols = {f"x{i}": rng.standard_normal(n) * scale for i in range(feats)}
    cols["y"] = sum(cols[c] for c in list(cols))
    return Frame(cols), [f"x{i}" for i in range(feats)]


def case_gate_large_scale(be):
    df, xs = _scaled_singular()                                                                         # :1309-1321
    assert be.eval(df, pds.lin_reg(*xs, target="y", add_bias=False)) is None
    df, xs = _scaled_full_rank()
    assert be.eval(df, pds.lin_reg(*xs, target="y", add_bias=False)) is not None


def case_gate_per_solver(be):
    for solver in ["qr", "svd", "choleskey"]:                                                           # :1324-1340
        df, xs = _scaled_singular()
        assert be.eval(df, pds.lin_reg(*xs, target="y", add_bias=False, solver=solver)) is None, solver
        df, xs = _scaled_full_rank()
        assert be.eval(df, pds.lin_reg(*xs, target="y", add_bias=False, solver=solver)) is not None, solver


# ---- null policies (tests/test_many.py:1636-1726) ----------------------------------------------------------
def _null_policy_frame():
    rng = np.random.default_rng(7)
    n = 200
    x1, x2, x3 = rng.random(n), rng.random(n), rng.random(n)
    y = 0.5 * x1 + 0.3 * x2 - 0.2 * x3 + rng.random(n) * 0.001
    null_rows = set(range(0, n, 10))
    x1n = [None if i in null_rows else float(v) for i, v in enumerate(x1)]
    return Frame({"x1": x1n, "x2": x2, "x3": x3, "y": y}), x1, np.array([i in null_rows for i in range(n)])


def case_null_policies(be):
    df, x1, isnull = _null_policy_frame()
    base = lambda d: be.eval(d, pds.lin_reg("x1", "x2", "x3", target="y"))
    atol = 1e-4 if _f32() else 1e-8
    got = be.eval(df, pds.lin_reg("x1", "x2", "x3", target="y", null_policy="skip"))
    assert np.allclose(got, base(df.filter(~isnull)), atol=atol)
    with pytest.raises(Exception):
        be.eval(df, pds.lin_reg("x1", "x2", "x3", target="y", null_policy="raise"))
    for pol, val in [("zero", 0.0), ("one", 1.0), ("0.5", 0.5)]:
        got = be.eval(df, pds.lin_reg("x1", "x2", "x3", target="y", null_policy=pol))
        filled = df.with_columns(x1=np.where(isnull, val, x1))
        assert np.allclose(got, base(filled), atol=atol), pol
    r = be.eval(df, pds.lin_reg("x1", "x2", "x3", target="y", null_policy="ignore", singular_x_tol=0.0))
    assert r is None or len(r) == 3
    with pytest.raises(Exception) as ei:
        be.eval(df, pds.lin_reg("x1", "x2", "x3", target="y", null_policy="not_a_policy"))
    msg = str(ei.value).lower()
    assert "invalid" in msg or "nullpolicy" in msg or "policy" in msg


def case_error_strings(be):
    """Error texts of series_to_mat_for_lr (linear_regression.rs:167,171,198)."""
    with pytest.raises(Exception, match="#Data < #features"):
        be.eval(Frame({"x1": [1.0], "x2": [2.0], "y": [1.0]}), pds.lin_reg("x1", "x2", target="y"))
    with pytest.raises(Exception, match="Nulls found"):
        be.eval(Frame({"x1": [1.0, None, 2.0], "y": [1.0, 2.0, 3.0]}), pds.lin_reg("x1", target="y", null_policy="raise"))


def case_weighted(be):
    """faer_weighted_lr (lr_solvers.rs:386-409) through pl_lr(weighted=True): vs the weighted normal equations."""
    rng = np.random.default_rng(12)
    n = 400
    X = rng.standard_normal((n, 3))
    y = X @ [0.3, -0.7, 1.1] + 0.2 + 0.1 * rng.standard_normal(n)
    w = rng.uniform(0.2, 2.0, n)
    df = Frame({"x1": X[:, 0], "x2": X[:, 1], "x3": X[:, 2], "y": y, "w": w})
    c = be.eval(df, pds.lin_reg("x1", "x2", "x3", target="y", weights="w", add_bias=True))
    Xb = np.column_stack([X, np.ones(n)])
    ref = np.linalg.solve((Xb.T * w) @ Xb, (Xb.T * w) @ y)
    np.testing.assert_allclose(c, ref, rtol=1e-4 if _f32() else 1e-9, atol=1e-5 if _f32() else 1e-11)


def case_int_and_chunked_inputs(be):
    """Features are not cast in Python (expr_linear.py:256-258): int columns, f32 columns in f64 mode, multi-chunk
    and sliced (offset != 0) inputs all go through the cast/packing path (utils/mod.rs:134-198)."""
    import pyarrow as pa

    rng = np.random.default_rng(13)
    n = 301
    xi = rng.integers(-50, 50, n)
    xf = rng.standard_normal(n).astype(np.float32)
    y = 0.25 * xi - 1.5 * xf + 3.0 + 0.01 * rng.standard_normal(n)
    xi_chunked = pa.chunked_array([pa.array(xi[:100]), pa.array(xi[100:])])
    full = Frame({"xi": xi_chunked, "xf": xf, "y": y})
    sliced = full.slice(7, 250)
    for fr, sl in ((full, slice(None)), (sliced, slice(7, 257))):
        c = be.eval(fr, pds.lin_reg("xi", "xf", target="y", add_bias=True))
        Xb = np.column_stack([xi[sl].astype(float), xf[sl].astype(float), np.ones(len(xi[sl]))])
        ref, *_ = np.linalg.lstsq(Xb, y[sl], rcond=None)
        np.testing.assert_allclose(c, ref, rtol=1e-3 if _f32() else 1e-9, atol=1e-4 if _f32() else 1e-10)


# ---------------------------------------------------------------------------------------------------------
def _logistic_case(be, bias, n):
    """test_logistic_reg_against_sklearn (test_linear_exprs.py:9-57): make_classification(10_000 x n, random_state=1)
    vs sklearn LogisticRegression(penalty=None, tol=1e-6, max_iter=400); the reference asserts the one-sided
    (ours - sklearn) < 1e-5, this restatement holds |ours - sklearn| to the accuracy sklearn's own tol=1e-6 stop
    leaves (a few 1e-6) and the predictions to 1e-5 two-sided."""
    import warnings

    from sklearn.datasets import make_classification
    from sklearn.linear_model import LogisticRegression

    X, y = make_classification(n_samples=10_000, n_features=n, n_redundant=0, n_informative=n - 1, random_state=1,
                               n_clusters_per_class=1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        clf = LogisticRegression(random_state=0, penalty=None, tol=1e-6, max_iter=400, fit_intercept=bias).fit(X, y)
    names = [f"x_{i}" for i in range(n)]
    df = Frame({nm: X[:, i] for i, nm in enumerate(names)} | {"y": y.astype(np.int64)})
    c = be.eval(df, pds.logistic_reg(*names, target="y", add_bias=bias, max_iter=400, tol=1e-6))
    test_tol = 1e-5
    if bias:
        assert np.all(np.abs(c[:-1] - clf.coef_[0]) < test_tol)
        assert abs(c[-1] - clf.intercept_[0]) < test_tol
    else:
        assert np.all(np.abs(c - clf.coef_[0]) < test_tol)
    pred, valid = be.eval(df, pds.logistic_reg(*names, target="y", add_bias=bias, max_iter=200, tol=1e-6, return_pred=True))
    assert valid.all()
    assert np.all(np.abs(pred - clf.predict_proba(X)[:, 1]) < test_tol)


def case_logistic_bias_5(be):
    _logistic_case(be, True, 5)


def case_logistic_bias_10(be):
    _logistic_case(be, True, 10)


def case_logistic_nobias_5(be):
    _logistic_case(be, False, 5)


def case_logistic_nobias_10(be):
    _logistic_case(be, False, 10)


def case_logistic_nulls_and_penalties(be):
    """Not in the reference's suite (it tests neither nulls nor l1 / l2 for logistic_reg): null_policy='skip' drops the
    rows and re-inserts nulls in the prediction column (logistic_regression.rs:74-93); l2 against sklearn's
    C = 1 / (m l2) (cost = mean log-loss + l2 / 2 |w|^2, logistic_solver.rs:54-70)."""
    import warnings

    from sklearn.linear_model import LogisticRegression

    rng = np.random.default_rng(77)
    m, p = 4000, 4
    X = rng.standard_normal((m, p))
    y = (rng.random(m) < 1.0 / (1.0 + np.exp(-(X @ np.array([1.0, -0.5, 0.25, 0.0]) + 0.3)))).astype(np.float64)
    names = [f"x{i}" for i in range(p)]
    drop = rng.random(m) < 0.05
    import pyarrow as pa

    cols = {nm: X[:, i] for i, nm in enumerate(names)} | {"y": y}
    cols["x1"] = pa.array(X[:, 1], mask=drop)
    df = Frame(cols)
    keep = ~drop
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        clf = LogisticRegression(penalty=None, tol=1e-10, max_iter=1000).fit(X[keep], y[keep])
    c = be.eval(df, pds.logistic_reg(*names, target="y", add_bias=True, tol=1e-8))
    assert np.max(np.abs(c - np.concatenate([clf.coef_[0], clf.intercept_]))) < 1e-6
    pred, valid = be.eval(df, pds.logistic_reg(*names, target="y", add_bias=True, tol=1e-8, return_pred=True))
    assert np.array_equal(valid, keep)
    assert np.max(np.abs(pred[keep] - clf.predict_proba(X[keep])[:, 1])) < 1e-6
    l2 = 0.05
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        clf2 = LogisticRegression(C=1.0 / (keep.sum() * l2), tol=1e-10, max_iter=1000).fit(X[keep], y[keep])
    c2 = be.eval(df, pds.logistic_reg(*names, target="y", add_bias=True, l2_reg=l2, tol=1e-8))
    assert np.max(np.abs(c2 - np.concatenate([clf2.coef_[0], clf2.intercept_]))) < 1e-6


ALL_CASES = [v for k, v in sorted(globals().items()) if k.startswith("case_") and callable(v)]
DUAL_DTYPE_CASES = {  # the reference runs these under both plugin variants (lin_reg_dtype fixture)
    "case_gate_collinear_nulls", "case_gate_off_finite", "case_gate_well_conditioned", "case_gate_group_by",
    "case_gate_return_pred", "case_gate_multi_target", "case_gate_large_scale", "case_gate_per_solver",
}
