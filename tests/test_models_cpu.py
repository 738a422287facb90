"""Model classes (`linear_models.py`) — host-side behaviour that needs no GPU: NaN policies (the reference's
tests/test_linear_models.py:9-49 restated with numpy), constructor / state errors with the reference's messages
(src/linear/mod.rs:20-31), and the loud failure of every numeric call when no CUDA device exists."""
import numpy as np
import pytest

from polars_ds_extension_b200.linear_models import LR, ElasticNet, OnlineLR, _df_to_xy, _handle_nans_in_np


def _data(size=5000, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.random((size, 3))
    nulls = x[:, 0] > 0.5
    x[nulls, 0] = np.nan
    y = (x[:, 0] + 0.2 * x[:, 1] - 0.3 * x[:, 2]).reshape(-1, 1)
    return x, y, nulls


def test_lr_null_policies_for_np():
    x, y, nulls = _data()
    x_nan, _ = _handle_nans_in_np(x, y, "ignore")
    assert np.all(np.isnan(x_nan[nulls][:, 0]))
    with pytest.raises(ValueError, match="Nulls found in X or y."):
        _handle_nans_in_np(x, y, "raise")
    x_skipped, y_skipped = _handle_nans_in_np(x, y, "skip")
    assert np.all(x_skipped == x[~nulls]) and len(y_skipped) == int((~nulls).sum())
    # the target is NaN wherever x1 is, so the fill policies drop those rows (null targets are always dropped)
    y_ok = np.where(nulls.reshape(-1, 1), 1.0, y)
    x_zeroed, _ = _handle_nans_in_np(x, y_ok, "zero")
    assert np.all(x_zeroed[nulls][:, 0] == 0.0)
    x_one, _ = _handle_nans_in_np(x, y_ok, "one")
    assert np.all(x_one[nulls][:, 0] == 1.0)
    x_num, _ = _handle_nans_in_np(x, y_ok, "1.25")
    assert np.all(x_num[nulls][:, 0] == 1.25)
    with pytest.raises(ValueError, match="Unknown null_policy"):
        _handle_nans_in_np(x, y_ok, "banana")
    with pytest.raises(ValueError, match="Unknown null_policy"):
        _handle_nans_in_np(x, y_ok, "inf")


def test_df_null_policies():
    from polars_ds_extension_b200 import Frame

    df = Frame({"a": [1.0, None, 3.0, 4.0], "b": [2.0, 1.0, None, 0.5], "y": [1.0, 2.0, 3.0, None]})
    X, y = _df_to_xy(df, ["a", "b"], "y", "skip")
    assert X.shape == (1, 2) and y.shape == (1, 1)
    X, y = _df_to_xy(df, ["a", "b"], "y", "zero")
    assert X.shape == (3, 2) and X[1, 0] == 0.0 and X[2, 1] == 0.0
    X, y = _df_to_xy(df, ["a", "b"], "y", "ignore")
    assert X.shape == (4, 2) and np.isnan(X[1, 0])
    with pytest.raises(ValueError, match="Nulls found in Dataframe."):
        _df_to_xy(df, ["a", "b"], "y", "raise")


def test_constructor_and_state_errors():
    with pytest.raises(ValueError, match="Cannot have both l1_reg and l2_reg <= 0."):
        ElasticNet(l1_reg=0.0, l2_reg=0.0)
    m = LR()
    assert not m.is_fit() and "Not fitted" in repr(m)
    with pytest.raises(ValueError, match="Matrix is not learned yet."):
        m.coeffs()
    with pytest.raises(ValueError, match="Matrix is not learned yet."):
        m.predict(np.zeros((2, 2)))
    o = OnlineLR()
    with pytest.raises(ValueError, match="You cannot update before the initial fit of the matrix."):
        o.update(np.zeros(3), 1.0)
    with pytest.raises(ValueError, match="Online regression currently must fit without null"):
        o.fit(np.array([[1.0, np.nan], [2.0, 3.0]]), np.array([1.0, 2.0]))
    with pytest.raises(ValueError, match="Not enough info to predict on a dataframe"):
        LR.from_values([1.0, 2.0]).predict_df({"a": np.zeros(2)})


def test_from_values_bias_rule():
    m = LR.from_values([1.0, -2.0], bias=0.5, feature_names_in_=["a", "b"])
    assert m.is_fit() and m.bias() == 0.5 and list(m.coeffs()) == [1.0, -2.0] and m.feature_names_in_ == ["a", "b"]
    m0 = LR.from_values([1.0, -2.0], bias=0.0)
    assert m0.bias() == 0.0 and m0.coeffs().size == 2
    e = ElasticNet.from_values([0.25], bias=1e-20)          # |bias| <= eps: no bias kept (lr_solvers.rs:113-114)
    assert not e.has_bias() and np.isnan(e.regularizers()[0])


def test_numeric_calls_fail_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    x, y, _ = _data(50)
    x = np.nan_to_num(x)
    y = np.nan_to_num(y)
    for call in (lambda: LR().fit(x, y), lambda: ElasticNet(0.1, 0.1).fit(x, y), lambda: OnlineLR().fit(x, y),
                 lambda: LR.from_values([1.0, 2.0, 3.0]).predict(x)):
        with pytest.raises(ValueError, match="no usable CUDA device"):
            call()
