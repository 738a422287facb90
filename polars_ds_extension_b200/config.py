"""Same switch as the reference (python/polars_ds/config.py:1,15-16)."""
LIN_REG_EXPR_F64 = True
"""If true, linear-regression expressions compute in f64 (`pl_lr*` symbols); if false in f32 (`pl_lr*_f32`)."""


def _which_lin_reg(x: str) -> str:
    return x if LIN_REG_EXPR_F64 else f"{x}_f32"
