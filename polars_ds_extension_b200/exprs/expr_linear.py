"""Host-side mirror of the reference's linear-regression expression builders.

Same names, arguments, defaults, validation and kwargs dictionaries as
/root/reference/python/polars_ds/exprs/expr_linear.py (lin_reg :105-274, lin_reg_w_rcond :356-410,
recursive_lin_reg :413-479, rolling_lin_reg :482-558, lin_reg_report :561-631): what crosses the plugin ABI — symbol
name, input order ([weights,] target(s), features...), which inputs are cast to the compute dtype, the pickled kwargs —
is identical, so the shared library sees the same calls Polars would make for the reference package.

Without a polars wheel the builders return a ``PluginExpr`` that ``Frame.select`` / ``Frame.group_by().agg`` evaluate
through the Polars-free harness; with polars installed ``PluginExpr.to_polars()`` gives the real ``pl.Expr``.
"""
from __future__ import annotations

import warnings
from typing import List, Literal, Union

from .. import config as cfg
from ..frame import ColExpr, PluginExpr, col
from ..typing import LRSolverMethods, NullPolicy

__all__ = ["lin_reg", "logistic_reg", "simple_lin_reg", "query_ar_coeffs", "linear_impute", "lin_reg_w_rcond", "recursive_lin_reg", "rolling_lin_reg", "lin_reg_report"]

ExprLike = Union[str, ColExpr]


def lr_formula(s: ExprLike) -> ColExpr:
    if isinstance(s, str):
        return col(s)
    if isinstance(s, ColExpr):
        return s
    raise ValueError("Input can only be a column name or a column expression.")


def _dtype() -> str:
    return "f64" if cfg.LIN_REG_EXPR_F64 else "f32"


def lin_reg(
    *x: ExprLike,
    target: Union[ExprLike, List[ExprLike]],
    add_bias: bool = False,
    weights: ExprLike | None = None,
    return_pred: bool = False,
    l1_reg: float = 0.0,
    l2_reg: float = 0.0,
    tol: float = 1e-5,
    solver: LRSolverMethods = "qr",
    max_iter: int = 200,
    null_policy: NullPolicy = "skip",
    positive: bool = False,
    singular_x_tol: float | None = None,
) -> PluginExpr:
    """Least squares / ridge / lasso / elastic net / non-negative fit of `target` on `x` (bias, if any, is last)."""
    dtype = _dtype()
    if singular_x_tol is None:
        singular_x_tol = 1e-12 if cfg.LIN_REG_EXPR_F64 else 1e-6

    if isinstance(target, list):
        n_targets = len(target)
        if n_targets == 0:
            raise ValueError("If `target` is a list, it cannot be empty.")
        if n_targets == 1:
            # the reference forwards everything except `positive` and `max_iter` (expr_linear.py:192-205)
            return lin_reg(*x, target=target[0], add_bias=add_bias, weights=weights, return_pred=return_pred,
                           l1_reg=l1_reg, l2_reg=l2_reg, tol=tol, solver=solver, null_policy=null_policy,
                           singular_x_tol=singular_x_tol)
        cols = [lr_formula(t).alias(f"target_{i}").cast(dtype) for i, t in enumerate(target)]
        kwargs = {"bias": add_bias, "null_policy": null_policy, "solver": solver, "last_target_idx": n_targets,
                  "l2_reg": l2_reg, "singular_x_tol": singular_x_tol}
        cols.extend(lr_formula(z) for z in x)
        if return_pred:
            return PluginExpr(cfg._which_lin_reg("pl_lr_multi_pred"), cols, kwargs, out_name="lr_pred")
        return PluginExpr(cfg._which_lin_reg("pl_lr_multi"), cols, kwargs, returns_scalar=True, out_name="coeffs")

    if max_iter <= 0:
        raise ValueError("Input `max_iter` must be a positive.")
    weighted = weights is not None
    kwargs = {"bias": add_bias, "null_policy": null_policy, "l1_reg": l1_reg, "l2_reg": l2_reg, "solver": solver,
              "tol": tol, "max_iter": max_iter, "weighted": weighted, "positive": positive,
              "singular_x_tol": singular_x_tol}
    if weighted:
        cols = [lr_formula(weights).cast(dtype).rechunk(), lr_formula(target).cast(dtype)]
    else:
        cols = [lr_formula(target).cast(dtype)]
    cols.extend(lr_formula(z) for z in x)
    if return_pred:
        return PluginExpr(cfg._which_lin_reg("pl_lr_pred"), cols, kwargs, out_name="lr_pred")
    return PluginExpr(cfg._which_lin_reg("pl_lr"), cols, kwargs, returns_scalar=True, out_name="coeffs")


def logistic_reg(*x: ExprLike, target: ExprLike, add_bias: bool = True, l1_reg: float = 0.0, l2_reg: float = 0.0,
                 tol: float = 1e-5, max_iter: int = 200, null_policy: NullPolicy = "skip",
                 return_pred: bool = False) -> PluginExpr:
    """Binary logistic regression (target must be 0 / 1): coefficients (bias last) or, with `return_pred`, the predicted
    probabilities.  Reference: expr_linear.py:277-354 -> pl_logistic_coeffs / pl_logistic_pred, always float64.  The
    reference minimises the mean log-loss (+ l2 / 2 |w|^2, + l1 |w|_1) with L-BFGS; here the same minimiser is reached
    by Newton / IRLS on the device (csrc/host/lr_host.cc::host_logistic)."""
    if max_iter <= 0:
        raise ValueError("Input `max_iter` must be a positive.")
    kwargs = {"bias": add_bias, "null_policy": null_policy, "l1_reg": l1_reg, "l2_reg": l2_reg, "solver": "",
              "tol": abs(tol), "max_iter": max_iter}
    cols = [lr_formula(target).cast("f64")]
    cols.extend(lr_formula(z) for z in x)
    if return_pred:
        return PluginExpr("pl_logistic_pred", cols, kwargs, out_name="__pred__")
    return PluginExpr("pl_logistic_coeffs", cols, kwargs, returns_scalar=True, out_name="__coeffs__")


def simple_lin_reg(x: ExprLike, target: ExprLike, add_bias: bool = False, weights: ExprLike | None = None,
                   return_pred: bool = False) -> PluginExpr:
    """One predictor, one target (reference: expr_linear.py:44-102, where the closed form beta = cov(x, y) / var(x),
    alpha = mean(y) - beta mean(x) is spelled out of Polars reductions, several passes over the columns).  Here it is
    the p = 1 case of the engine: one pass builds the 3 x 3 moments, the 1 x 1 / 2 x 2 solve is the same closed form.
    No rank gate (a constant x gives what 0 / 0 gives in the closed form: no finite coefficients)."""
    return lin_reg(x, target=target, add_bias=add_bias, weights=weights, return_pred=return_pred, singular_x_tol=0.0)


def lin_reg_w_rcond(*x: ExprLike, target: ExprLike, add_bias: bool = False, rcond: float = 0.0, l2_reg: float = 0.0,
                    null_policy: NullPolicy = "raise") -> PluginExpr:
    """SVD-based least squares that zeroes small singular values; returns coefficients and the singular values of X."""
    cols = [lr_formula(target).cast(_dtype())]
    cols.extend(lr_formula(z) for z in x)
    kwargs = {"bias": add_bias, "null_policy": null_policy, "l1_reg": 0.0, "l2_reg": l2_reg, "solver": "",
              "tol": abs(rcond)}
    return PluginExpr(cfg._which_lin_reg("pl_lr_w_rcond"), cols, kwargs)


def recursive_lin_reg(*x: ExprLike, target: ExprLike, start_with: int, add_bias: bool = False, l2_reg: float = 0.0,
                      null_policy: NullPolicy = "raise") -> PluginExpr:
    """Expanding-window least squares: row j holds the fit on rows [0, j]; the first `start_with`-1 rows are null."""
    if start_with < 1:
        raise ValueError("You must start with >= 1 rows for recursive linear regression.")
    cols = [lr_formula(target).cast(_dtype())]
    features = [lr_formula(z) for z in x]
    if len(features) > start_with:
        warnings.warn("# features > number of rows for the initial fit. Outputs may be off.", stacklevel=2)
    cols.extend(features)
    kwargs = {"null_policy": null_policy, "n": start_with, "bias": add_bias, "lambda": abs(l2_reg), "min_size": 0}
    return PluginExpr(cfg._which_lin_reg("pl_recursive_lr"), cols, kwargs)


def rolling_lin_reg(*x: ExprLike, target: ExprLike, window_size: int, add_bias: bool = False, l2_reg: float = 0.0,
                    min_valid_rows: int | None = None, null_policy: NullPolicy = "raise") -> PluginExpr:
    """Rolling-window least squares: row j holds the fit on rows (j - window_size, j]."""
    if window_size < 2:
        raise ValueError("`window_size` must be >= 2.")
    cols = [lr_formula(target).cast(_dtype())]
    features = [lr_formula(z) for z in x]
    if len(features) > window_size:
        raise ValueError("# features > window size. Linear regression is not well-defined.")
    if min_valid_rows is None:
        min_size = min(len(features), window_size)
    else:
        if min_valid_rows < len(features):
            warnings.warn("# features > min_window_size. Linear regression may not always be well-defined.",
                          stacklevel=2)
        min_size = min_valid_rows
    cols.extend(features)
    kwargs = {"null_policy": null_policy, "n": window_size, "bias": add_bias, "lambda": abs(l2_reg),
              "min_size": min_size}
    return PluginExpr(cfg._which_lin_reg("pl_rolling_lr"), cols, kwargs)


def lin_reg_report(*x: ExprLike, target: ExprLike, weights: ExprLike | None = None, add_bias: bool = False,
                   null_policy: NullPolicy = "raise",
                   std_err: Literal["se", "hc0", "hc1", "hc2", "hc3"] = "se") -> PluginExpr:
    """OLS (or WLS) report: beta, standard errors (classic or HC0-HC3), t, p, 95% CI, r2, adjusted r2 per coefficient."""
    kwargs = {"bias": add_bias, "null_policy": null_policy, "l1_reg": 0.0, "l2_reg": 0.0, "solver": "qr", "tol": 0.0,
              "std_err": std_err.lower()}
    dtype = _dtype()
    t = lr_formula(target).cast(dtype)
    if weights is None:
        cols = [t.var(), t]
        symbol = cfg._which_lin_reg("pl_lin_reg_report")
    else:
        cols = [lr_formula(weights).cast(dtype).rechunk(), t.var(), t]
        symbol = cfg._which_lin_reg("pl_wls_report")
    cols.extend(lr_formula(z) for z in x)
    return PluginExpr(symbol, cols, kwargs, changes_length=True)


# ---------------------------------------------------------------------------------------------------------------
# callers that funnel into lin_reg
# ---------------------------------------------------------------------------------------------------------------
def query_ar_coeffs(x: ExprLike, lag: int, add_bias: bool = True, null_policy: NullPolicy = "raise") -> PluginExpr:
    """Autoregressive coefficients of order `lag` (reference: exprs/ts_features.py:419-461): x_t regressed on
    x_{t-1} .. x_{t-lag} over rows lag.., bias (if any) last.  One lin_reg call, hence one pass of the moments kernel."""
    if null_policy not in ("raise", "one", "zero"):
        try:
            import math

            if not math.isfinite(float(null_policy)):
                raise ValueError
        except Exception:
            raise ValueError(
                "`null_polocy` must be 'raise', 'one', 'zero' or any finite numeric string for AR coefficients."
            )
    if lag <= 0:
        raise ValueError("`lag` must be > 0.")
    xx = lr_formula(x)
    return lin_reg(*[xx.shift(i).slice(lag).alias(str(i)) for i in range(1, lag + 1)], target=xx.slice(lag),
                   add_bias=add_bias, null_policy=null_policy)


def linear_impute(df, features: List[str], target: str, add_bias: bool = False):
    """Fill the nulls of `target` with the prediction of a linear regression on `features`, fitted on the rows where
    everything is present (reference: pipeline/transforms.py:112-155; null_policy="skip", target cast to f64).
    Returns the frame with the imputed column; rows whose features are null stay null, as `sum_horizontal` of the
    reference propagates nothing there either (Polars' sum_horizontal skips nulls: those rows get the partial sum)."""
    import numpy as np
    import pyarrow as pa

    coeffs = np.asarray(df.select(lin_reg(*features, target=target, add_bias=add_bias, null_policy="skip"))["coeffs"][0].as_py(),
                        dtype=np.float64)
    y = df[target].cast(pa.float64()).combine_chunks()
    missing = ~np.asarray(y.is_valid().to_numpy(zero_copy_only=False), dtype=bool)
    pred = np.full(len(y), coeffs[-1] if add_bias else 0.0)
    for f, b in zip(features, coeffs):
        col_f = df[f].cast(pa.float64()).combine_chunks()
        ok = np.asarray(col_f.is_valid().to_numpy(zero_copy_only=False), dtype=bool)
        pred += np.where(ok, col_f.fill_null(0.0).to_numpy(zero_copy_only=False), 0.0) * b   # sum_horizontal ignores nulls
    vals = np.where(missing, pred, y.fill_null(0.0).to_numpy(zero_copy_only=False))
    return df.with_columns(**{target: pa.array(vals)})
