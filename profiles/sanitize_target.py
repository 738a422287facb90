"""compute-sanitizer target: every kernel family once, at small sizes, through the plugin C ABI / device layer.

    compute-sanitizer --tool memcheck  python profiles/sanitize_target.py > profiles/sanitizer_memcheck_r02.log
    compute-sanitizer --tool racecheck python profiles/sanitize_target.py > profiles/sanitizer_racecheck_r02.log
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import polars_ds_extension_b200 as pds  # noqa: E402
import polars_ds_extension_b200.config as cfg  # noqa: E402
from polars_ds_extension_b200 import device as dev  # noqa: E402
from polars_ds_extension_b200._lib import lib  # noqa: E402
from tests.backends import PluginBackend  # noqa: E402

gpu = PluginBackend()
rng = np.random.default_rng(0)


def frame(n, p, dt):
    X = rng.standard_normal((n, p)).astype(dt)
    y = (X @ (((np.arange(p) % 7) - 3) / 4.0) + 0.1 * rng.standard_normal(n)).astype(dt)
    return pds.Frame({f"x{i}": X[:, i] for i in range(p)} | {"y": y}), [f"x{i}" for i in range(p)]


for f64 in (False, True):
    cfg.LIN_REG_EXPR_F64 = f64
    dt = np.float64 if f64 else np.float32
    df, xs = frame(9000, 32, dt)                       # tcgen05 kernel with side warps (f32) / DMMA (f64)
    gpu.eval(df, pds.lin_reg(*xs, target="y", return_pred=True))
    df, xs = frame(8200, 64, dt)                       # tcgen05 kernel, 4 quadrants, converter-side sums (f32)
    gpu.eval(df, pds.lin_reg(*xs, target="y", add_bias=True))
    df, xs = frame(5000, 8, dt)
    gpu.eval(df, pds.rolling_lin_reg(*xs, target="y", window_size=1024))       # packed f32x2 / lane-per-moment
    gpu.eval(df, pds.rolling_lin_reg(*xs[:3], target="y", window_size=7, add_bias=True))
    gpu.eval(df, pds.recursive_lin_reg(*xs, target="y", start_with=9, add_bias=True))
    df14, xs14 = frame(2100, 14, dt)
    gpu.eval(df14, pds.rolling_lin_reg(*xs14, target="y", window_size=64))      # generic path
    gpu.eval(df, pds.lin_reg_report(*xs, target="y", add_bias=True, std_err="hc3"))
    gpu.eval(df, pds.lin_reg(*xs, target="y", l1_reg=0.01, add_bias=True))
    gpu.eval(df, pds.lin_reg(*xs, target="y", positive=True))
    gpu.eval(df, pds.lin_reg(*xs, target=["y", "x0", "x1"], add_bias=True))        # multi-target: side sums for t > 1
    if f64:
        yb = (rng.random(5000) < 0.5).astype(np.float64)
        gpu.eval(df.with_columns(yb=yb), pds.logistic_reg(*xs, target="yb", max_iter=5))       # K11 + weighted K2a + K3
    gid = np.repeat(np.arange(10), 500)
    gpu.group_eval(df.with_columns(g=gid), "g", pds.lin_reg(*xs, target="y", add_bias=True), fast=True)
    gpu.group_eval(df.with_columns(g=gid), "g", pds.lin_reg(*xs, target="y", l1_reg=0.01), fast=True)
# device layer: the DMMA ring kernels (side / wide, every block-count class, weights / mask, n % 8 != 0), f64 and f32 columns
for tdt in (torch.float64, torch.float32):
    for p_, t_, wt, mk in [(8, 1, False, False), (32, 1, True, True), (30, 2, True, False), (64, 1, False, True), (62, 1, False, False),
                           (16, 1, True, False), (46, 1, False, False)]:
        n_, ld_ = 4099, 4128
        Zd = torch.randn((p_ + t_, ld_), device="cuda", dtype=tdt)
        w_ = (torch.rand(ld_, device="cuda", dtype=tdt) + 0.5) if wt else None
        m_ = (torch.rand(ld_, device="cuda") > 0.2).to(tdt) if mk else None
        if tdt == torch.float32:
            lib().pdsb_set_moments_path(1)
        dev.moments(Zd[:p_], Zd[p_:], n=n_, w=w_, mask=m_)
        lib().pdsb_set_moments_path(0)
# grouped path with several items per group (work queue, precomputed item rows, shared-memory solve)
cfg.LIN_REG_EXPR_F64 = False
dfg, xsg = frame(27000, 8, np.float32)
gidg = np.repeat(np.arange(3), 9000)
gpu.group_eval(dfg.with_columns(g=gidg), "g", pds.lin_reg(*xsg, target="y", add_bias=True), fast=True)
# device layer: frames
Z = torch.randn((33, 8192), device="cuda")
fr = dev.to_frame(Z, n=8192)
dev.moments_frame(fr, 8192, 33, 0, 32, 32, 1)
torch.cuda.synchronize()
print("sanitize target done; kernels launched:", lib().pdsb_kernel_launch_count())
