"""Writes tests/golden/lin_reg_golden.npz: frozen inputs + expected outputs for the lin_reg family.

The reference (a Rust cdylib) can neither be built nor imported in this environment, so the expected values are NOT
produced by it — and deliberately not by this repository's oracle either.  They come from the external checkers the
reference's own tests compare against (scikit-learn, numpy.linalg, scipy: /root/reference/tests/test_linear_exprs.py
:61-1340, tests/test_linear_models.py:52-160), evaluated once here on seeded inputs and frozen.  Both the CPU oracle
and the CUDA path are then tested against the same file (tests/test_golden_vectors.py).

    python tests/golden/make_golden.py      # numpy 2.3, scikit-learn and scipy as in the image
"""
from pathlib import Path

import numpy as np
from scipy import optimize, stats
from sklearn import linear_model

out = {}
rng = np.random.default_rng(20260922)
n = 1500
X = rng.random((n, 3))
y = X @ [0.5, 0.1, -0.15] + 0.3 + 0.05 * rng.standard_normal(n)
w = rng.random(n) + 0.25
out["X"], out["y"], out["w"] = X, y, w
A = np.column_stack([X, np.ones(n)])

# pds.lin_reg, add_bias=True                                  (test_lin_reg_against_sklearn, :61-120)
r = linear_model.LinearRegression(fit_intercept=True).fit(X, y)
out["ols_bias"] = np.append(r.coef_, r.intercept_)
# no bias
out["ols"] = np.linalg.lstsq(X, y, rcond=None)[0]
# ridge, bias unpenalised                                      (same test, Ridge(alpha=0.1))
r = linear_model.Ridge(alpha=0.1, fit_intercept=True, tol=1e-12, solver="cholesky").fit(X, y)
out["ridge_bias"] = np.append(r.coef_, r.intercept_)
# lasso / elastic net: alpha, l1_ratio as the reference maps them (test_lasso :640-700, test_elastic_net :700-716)
r = linear_model.Lasso(alpha=0.01, fit_intercept=True, tol=1e-10, max_iter=100000).fit(X, y)
out["lasso_bias"] = np.append(r.coef_, r.intercept_)
l1, l2 = 0.01, 0.02
r = linear_model.ElasticNet(alpha=l1 + l2, l1_ratio=l1 / (l1 + l2), fit_intercept=False, tol=1e-10, max_iter=100000).fit(X, y)
out["enet"] = r.coef_
# non-negative least squares                                    (test_positive_lin_reg :600-640 uses sklearn positive=True)
Xn = X.copy(); Xn[:, 2] = -Xn[:, 2] * 0.5
out["Xn"] = Xn
out["nnls"] = optimize.nnls(Xn, y)[0]
# weighted least squares                                       (test_wls :1260-1300)
sw = np.sqrt(w)
out["wls_bias"] = np.linalg.lstsq(A * sw[:, None], y * sw, rcond=None)[0]
# lin_reg_w_rcond: numpy lstsq coefficients + singular values   (test_lstsq_w_rcond :550-600)
c, _, _, sv = np.linalg.lstsq(X, y, rcond=0.0)
out["rcond_coeffs"], out["rcond_sv"] = c, sv
# multi-target                                                  (test_multi_target :440-520)
y2 = X @ [-0.2, 0.7, 0.05] + 0.02 * rng.standard_normal(n)
out["y2"] = y2
out["multi"] = np.linalg.lstsq(A, np.column_stack([y, y2]), rcond=None)[0].T      # [target][coef]
# report (se / t / p / CI)                                      (test_lin_reg_report :984-1028)
beta = out["ols_bias"]
e = y - A @ beta
dof = n - 4
se = np.sqrt(np.diag(np.linalg.inv(A.T @ A)) * (e @ e) / dof)
t = beta / se
out["report"] = np.vstack([beta, se, t, 2 * stats.t.sf(np.abs(t), dof), beta - stats.t.ppf(0.975, dof) * se,
                           beta + stats.t.ppf(0.975, dof) * se])
# rolling (window OLS) and recursive (prefix OLS) on the first 400 rows   (:123-166, 718-854)
m, win, start = 400, 25, 10
roll = np.full((m, 3), np.nan)
for j in range(win - 1, m):
    roll[j] = np.linalg.lstsq(X[j - win + 1: j + 1], y[j - win + 1: j + 1], rcond=None)[0]
rec = np.full((m, 3), np.nan)
for j in range(start - 1, m):
    rec[j] = np.linalg.lstsq(X[: j + 1], y[: j + 1], rcond=None)[0]
out["rolling_w25"], out["recursive_s10"] = roll, rec
# group_by: per-group OLS with bias on 6 uneven groups
keys = np.repeat(np.arange(6), [40, 250, 7, 500, 300, 403])
out["keys"] = keys
out["grouped_bias"] = np.vstack([np.linalg.lstsq(A[keys == g], y[keys == g], rcond=None)[0] for g in range(6)])

np.savez_compressed(Path(__file__).with_name("lin_reg_golden.npz"), **out)
print({k: v.shape for k, v in out.items()})
