"""Callers that funnel into `lin_reg`: `query_ar_coeffs` (reference exprs/ts_features.py:419-461)
and `linear_impute` (pipeline/transforms.py:112-155).  The reference has no numeric test for either; the expected values
are the definitions (numpy.linalg.lstsq on the lagged design / on the complete rows).  CPU: through the oracle backend;
GPU: through the plugin C ABI."""
import numpy as np
import pyarrow as pa
import pytest

import polars_ds_extension_b200 as pds
from polars_ds_extension_b200 import Frame


def _series(n=3000, seed=5):
    rng = np.random.default_rng(seed)
    x = np.zeros(n)
    e = rng.standard_normal(n)
    for t in range(3, n):
        x[t] = 0.4 + 0.5 * x[t - 1] - 0.25 * x[t - 2] + 0.1 * x[t - 3] + e[t]
    return x


def _ar_expected(x, lag, add_bias):
    cols = [x[lag - i: len(x) - i] for i in range(1, lag + 1)]
    A = np.column_stack(cols + ([np.ones(len(x) - lag)] if add_bias else []))
    return np.linalg.lstsq(A, x[lag:], rcond=None)[0]


def _check_ar(be):
    x = _series()
    df = Frame({"x": x})
    for lag, bias in [(1, True), (3, True), (5, False)]:
        got = be.eval(df, pds.query_ar_coeffs("x", lag=lag, add_bias=bias))
        np.testing.assert_allclose(got, _ar_expected(x, lag, bias), rtol=1e-7, atol=1e-9)
    with pytest.raises(ValueError, match="`lag` must be > 0."):
        pds.query_ar_coeffs("x", lag=0)
    with pytest.raises(ValueError, match="must be 'raise', 'one', 'zero'"):
        pds.query_ar_coeffs("x", lag=2, null_policy="skip")
    assert pds.query_ar_coeffs("x", lag=2, null_policy="0.5").kwargs["null_policy"] == "0.5"


def _impute_frame():
    rng = np.random.default_rng(8)
    n = 2000
    a, b = rng.standard_normal(n), rng.standard_normal(n)
    y = 1.5 * a - 0.5 * b + 0.25 + 0.01 * rng.standard_normal(n)
    ym = y.copy()
    miss = rng.random(n) < 0.2
    ym[miss] = np.nan
    df = Frame({"a": a, "b": b, "y": pa.array(ym, mask=miss)})
    A = np.column_stack([a, b, np.ones(n)])
    beta = np.linalg.lstsq(A[~miss], y[~miss], rcond=None)[0]
    return df, miss, y, A @ beta


def _check_impute(df, miss, y, pred):
    out = pds.linear_impute(df, ["a", "b"], "y", add_bias=True)
    got = out["y"].to_numpy()
    assert out["y"].null_count == 0
    np.testing.assert_array_equal(got[~miss], y[~miss])
    np.testing.assert_allclose(got[miss], pred[miss], rtol=1e-8, atol=1e-10)


def test_ar_coeffs_oracle():
    from tests.backends import OracleBackend

    _check_ar(OracleBackend())


def test_linear_impute_oracle(monkeypatch):
    from tests.backends import OracleBackend

    be = OracleBackend()
    monkeypatch.setattr(Frame, "evaluate", lambda self, e: pa.array([list(be.eval(self, e))]))
    _check_impute(*_impute_frame())


@pytest.mark.gpu
def test_ar_coeffs_cuda():
    from tests.backends import PluginBackend

    _check_ar(PluginBackend())


@pytest.mark.gpu
def test_linear_impute_cuda():
    _check_impute(*_impute_frame())
