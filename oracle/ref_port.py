"""Python binding of oracle/ref_port.c — TEST / BASELINE INFRASTRUCTURE (never imported by the product).

`lr_pred_f32` = the reference's `pl_lr_pred_f32` on its no-null fast path
(/root/reference/src/num_ext/linear_regression_f32.rs:568-685): pack -> Gram -> gated QR solve -> predict -> resid ->
output copies, with the reference's thread structure per phase (see ref_port.c).  The q x q solve + rank gate is the
numpy oracle's own restatement of faer_solve_lr_gated (lin_reg_oracle.py), applied to the C-built Gram.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from . import lin_reg_oracle as orc

_DIR = Path(__file__).resolve().parent
_SO = _DIR / "_ref" / "libref_port.so"
_lib = None


def build() -> Path:
    """gcc -O3 -fopenmp oracle/ref_port.c -> oracle/_ref/libref_port.so (recipe: oracle/Makefile)."""
    subprocess.run(["make", "-s", "-C", str(_DIR)], check=True)
    return _SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not _SO.exists():
            build()
        L = C.CDLL(str(_SO))
        L.ref_pack_f32.restype = C.c_void_p
        L.ref_pack_f32.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int64, C.c_int]
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_gram_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_predict_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_now.restype = C.c_double
        L.ref_max_threads.restype = C.c_int
        _lib = L
    return _lib


def set_threads(n: int) -> None:
    """torchrun exports OMP_NUM_THREADS=1; the baseline must say how many threads it really used."""
    lib().ref_set_threads(int(n))


def threads() -> int:
    return int(lib().ref_max_threads())


def lr_pred_f32(cols, add_bias: bool = False, singular_x_tol: float = 1e-6, l2_reg: float = 0.0, solver: str = "qr"):
    """cols = [y, x_1, ..., x_p] as contiguous float32 arrays.  Returns (coeffs | None, pred, resid, phase seconds)."""
    L = lib()
    p = len(cols) - 1
    n = len(cols[0])
    for c in cols:
        assert c.dtype == np.float32 and c.flags.c_contiguous and len(c) == n
    ptrs = (C.c_void_p * (p + 1))(*[c.ctypes.data for c in cols])
    t0 = L.ref_now()
    P = L.ref_pack_f32(ptrs, p, n, int(add_bias))
    if not P:
        raise MemoryError("ref_pack_f32")
    t1 = L.ref_now()
    try:
        q = p + int(add_bias)
        xtx = np.empty((q, q), dtype=np.float32)
        xty = np.empty(q, dtype=np.float32)
        L.ref_gram_f32(P, xtx.ctypes.data, xty.ctypes.data)
        t2 = L.ref_now()
        beta = _gated_solve(xtx, xty.reshape(-1, 1), np.float32(l2_reg), add_bias, solver, np.float32(singular_x_tol))
        t3 = L.ref_now()
        pred = np.empty(n, dtype=np.float32)
        resid = np.empty(n, dtype=np.float32)
        if beta is not None:
            b = np.ascontiguousarray(beta[:, 0], dtype=np.float32)
            L.ref_predict_f32(P, b.ctypes.data, pred.ctypes.data, resid.ctypes.data)
        t4 = L.ref_now()
    finally:
        L.ref_free(P)
    times = {"pack_1thread": t1 - t0, "gram_xty_all_threads": t2 - t1, "solve": t3 - t2,
             "predict_all_threads_resid_copy_1thread": t4 - t3, "total": t4 - t0}
    return (None if beta is None else beta[:, 0]), pred, resid, times


def _gated_solve(xtx, xty, lam, add_bias, how, tol):
    """faer_solve_lr_gated (lr_solvers.rs:329-382) / faer_solve_lr (:299-308) on an already-built Gram."""
    dt = xtx.dtype.type
    q = xtx.shape[0]
    n1 = q - int(add_bias)
    if lam > 0 and n1 >= 1:
        idx = np.arange(n1)
        xtx[idx, idx] += dt(lam)
    if not tol > 0:
        return orc.solve_xtx_xty(xtx, xty, how)
    d = np.diag(xtx)
    if np.any(d <= 0):
        return None
    ln_den = dt(np.sum(np.log(d), dtype=dt))
    qrp = orc._qr_piv(xtx)
    with np.errstate(divide="ignore"):
        ln_det = dt(np.sum(np.log(np.abs(np.diag(qrp[1]))), dtype=dt))
    if ln_det - ln_den <= dt(np.log(dt(tol))):
        return None
    return orc._qr_solve(qrp, xty)
