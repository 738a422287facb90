"""Model classes on the GPU.  The first three tests restate the reference's own known-answer tests
(/root/reference/tests/test_linear_models.py:52-160: LR per solver, OnlineLR fit + 10 updates, ElasticNet with and
without bias — same shapes, same generating model, same scikit-learn checkers, same tolerances); the unseeded
`pds.random` frames are re-created with a seeded numpy generator.  The rest pins layouts, errors and the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _frame(size=5000, seed=0, noise=1e-4):
    rng = np.random.default_rng(seed)
    X = rng.random((size, 3))
    y = X[:, 0] + 0.2 * X[:, 1] - 0.3 * X[:, 2] + noise * rng.random(size)
    return X, y.reshape(-1, 1)


@pytest.mark.parametrize("solver", ["svd", "cholesky", "qr", "choleskey"])
def test_lr(solver):
    from sklearn.linear_model import LinearRegression

    from polars_ds_extension_b200.linear_models import LR

    X, y = _frame()
    ols = LR(False, 0.0, solver).fit(X, y)
    sk = LinearRegression(fit_intercept=False).fit(X, y)
    assert np.all(np.abs(ols.coeffs() - sk.coef_) < 1e-6)
    np.testing.assert_allclose(ols.predict(X), sk.predict(X).reshape(-1, 1), atol=1e-6)


def test_lr_bias_ridge_layouts_and_null_policy():
    from sklearn.linear_model import Ridge

    from polars_ds_extension_b200.linear_models import LR

    X, y = _frame(seed=3, noise=0.05)
    y = y + 0.7
    ref = Ridge(alpha=0.5, fit_intercept=True).fit(X, y.ravel())
    # sklearn's Ridge with intercept does not penalise the intercept, exactly like lambda on the non-bias diagonal
    for Xv in (X, np.asfortranarray(X), np.column_stack([X, X])[:, :3], X[::1]):
        m = LR(has_bias=True, lambda_=0.5).fit(Xv, y)
        np.testing.assert_allclose(m.coeffs(), ref.coef_, atol=1e-8)
        assert abs(m.bias() - ref.intercept_) < 1e-8
        np.testing.assert_allclose(m.predict(Xv).ravel(), ref.predict(X), atol=1e-8)
    Xn = X.copy()
    Xn[::7, 1] = np.nan
    keep = ~np.isnan(Xn[:, 1])
    ref2 = Ridge(alpha=0.5, fit_intercept=True).fit(X[keep], y.ravel()[keep])
    m = LR(has_bias=True, lambda_=0.5).fit(Xn, y, null_policy="skip")
    np.testing.assert_allclose(m.coeffs(), ref2.coef_, atol=1e-8)
    with pytest.raises(ValueError, match="Nulls found in X or y."):
        LR().fit(Xn, y, null_policy="raise")


def test_online_lr():
    from sklearn.linear_model import LinearRegression

    from polars_ds_extension_b200.linear_models import OnlineLR

    X, y = _frame(seed=1)
    olr = OnlineLR()
    olr.fit(X[:10], y[:10])
    sk = LinearRegression(fit_intercept=False).fit(X[:10], y[:10])
    assert np.all(np.abs(olr.predict(X[:10]).flatten() - sk.predict(X[:10]).flatten()) < 1e-6)
    assert np.all(np.abs(olr.coeffs() - sk.coef_) < 1e-6)
    for i in range(10, 20):
        olr.update(X[i], y[i])
        sk = LinearRegression(fit_intercept=False).fit(X[: i + 1], y[: i + 1])
        assert np.all(np.abs(olr.coeffs() - sk.coef_) < 1e-6)
    # a NaN row is ignored; removing a row (c = -1) undoes its update
    before, inv_before = olr.coeffs(), olr.inv()
    olr.update(np.array([np.nan, 1.0, 2.0]), 3.0)
    np.testing.assert_array_equal(olr.coeffs(), before)
    olr.update(X[30], y[30]).update(X[30], y[30], c=-1.0)
    np.testing.assert_allclose(olr.coeffs(), before, atol=1e-9)
    np.testing.assert_allclose(olr.inv(), inv_before, rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(olr.inv(), np.linalg.inv(X[:20].T @ X[:20]), rtol=1e-6)


def test_online_lr_bias_ridge_and_restore():
    from polars_ds_extension_b200.linear_models import OnlineLR

    X, y = _frame(seed=2, noise=0.01)
    y = y + 0.3
    lam = 0.1
    olr = OnlineLR(lambda_=lam, has_bias=True).fit(X[:50], y[:50])
    for i in range(50, 80):
        olr.update(X[i], float(y[i, 0]))
    A = np.column_stack([X[:80], np.ones(80)])
    G = A.T @ A + np.diag([lam, lam, lam, 0.0])          # faer_qr_lr_with_inv: lambda on the non-bias diagonal
    w = np.linalg.solve(G, A.T @ y[:80]).ravel()
    np.testing.assert_allclose(np.append(olr.coeffs(), olr.bias()), w, atol=1e-8)
    clone = OnlineLR.from_coeffs_bias_inverse(olr.coeffs(), olr.bias(), olr.inv())
    clone.update(X[90], y[90])
    olr.update(X[90], y[90])
    np.testing.assert_allclose(clone.coeffs(), olr.coeffs(), atol=1e-12)
    with pytest.raises(ValueError, match="Dimension mismatch."):
        olr.update(np.zeros(5), 1.0)


@pytest.mark.parametrize("add_bias", [False, True])
def test_elastic_net(add_bias):
    import sklearn.linear_model as lm

    from polars_ds_extension_b200.linear_models import ElasticNet

    l1_reg = l2_reg = 0.1
    X, y = _frame(seed=4, noise=0.0)
    en = ElasticNet(l1_reg=l1_reg, l2_reg=l2_reg, has_bias=add_bias).fit(X, y)
    sk = lm.ElasticNet(alpha=l1_reg + l2_reg, l1_ratio=l1_reg / (l1_reg + l2_reg), fit_intercept=add_bias).fit(X, y)
    assert np.all(np.abs(en.coeffs() - sk.coef_) < 1e-4)
    if add_bias:
        assert abs(en.bias() - sk.intercept_) < 1e-4


def test_models_match_the_oracle_and_report_errors():
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    from oracle import lin_reg_oracle as orc
    from polars_ds_extension_b200 import Frame
    from polars_ds_extension_b200.linear_models import LR, ElasticNet

    X, y = _frame(size=2000, seed=9, noise=0.02)
    w = orc.faer_solve_lr(np.column_stack([X, np.ones(len(X))]), y, 0.25, True, "qr").ravel()
    m = LR(has_bias=True, lambda_=0.25).fit(X, y)
    np.testing.assert_allclose(np.append(m.coeffs(), m.bias()), w, atol=1e-9)
    cd = orc.faer_coordinate_descent(X, y, 0.01, 0.02, False, 1e-7, 5000, False).ravel()
    e = ElasticNet(0.01, 0.02, tol=1e-7, max_iter=5000).fit(X, y)
    np.testing.assert_allclose(e.coeffs(), cd, atol=1e-6)
    with pytest.raises(ValueError, match="Dimension mismatch."):
        LR().fit(X, y[:-1])
    with pytest.raises(ValueError, match="Not enough rows / columns."):
        LR().fit(X[:2], y[:2])
    with pytest.raises(ValueError, match="Dimension mismatch."):
        m.predict(X[:, :2])
    df = Frame({"a": X[:, 0], "b": X[:, 1], "c": X[:, 2], "y": y.ravel()})
    f = LR(has_bias=True, lambda_=0.25).fit_df(df, ["a", "b", "c"], "y")
    np.testing.assert_allclose(f.coeffs(), m.coeffs(), atol=1e-12)
    out = f.predict_df(df, name="yhat")
    np.testing.assert_allclose(np.asarray(out["yhat"].to_numpy()), m.predict(X).ravel(), atol=1e-12)


@pytest.mark.parametrize("add_bias", [False, True])
@pytest.mark.parametrize("weighted", [False, True])
def test_simple_lin_reg_is_the_closed_form(add_bias, weighted):
    """`pds.simple_lin_reg` (expr_linear.py:44-102): beta = cov_w(x, y) / var_w(x), alpha = mean_w(y) - beta mean_w(x)."""
    import polars_ds_extension_b200 as pds

    rng = np.random.default_rng(21)
    n = 20_000
    x = rng.standard_normal(n) * 2.0 + 1.0
    y = 0.75 * x - 0.4 + 0.3 * rng.standard_normal(n)
    w = rng.random(n) + 0.1
    ww = w if weighted else np.ones(n)
    if add_bias:
        xm, ym = np.sum(ww * x) / ww.sum(), np.sum(ww * y) / ww.sum()
        beta = np.sum(ww * (x - xm) * (y - ym)) / np.sum(ww * (x - xm) ** 2)
        want = [beta, ym - beta * xm]
    else:
        want = [np.sum(ww * x * y) / np.sum(ww * x * x)]
    df = pds.Frame({"x": x, "y": y, "w": w})
    e = pds.simple_lin_reg("x", "y", add_bias=add_bias, weights="w" if weighted else None)
    got = df.select(e)["coeffs"][0].as_py()
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-11)
    pr = df.select(pds.simple_lin_reg("x", "y", add_bias=add_bias, weights="w" if weighted else None, return_pred=True))
    arr = pr["lr_pred"]
    arr = arr.combine_chunks() if hasattr(arr, "combine_chunks") else arr
    pred = np.asarray(arr.field("pred").to_numpy(zero_copy_only=False))
    np.testing.assert_allclose(pred, want[0] * x + (want[1] if add_bias else 0.0), rtol=1e-9, atol=1e-9)
