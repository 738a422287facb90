"""The f32 acceptance rule of SURVEY.md §8c, in one place.

north_star: results must match the reference to 1e-4 relative for f32 (1e-6 for f64).  The reference's f32 twin is
restated by the oracle in f32 mode (`o32`); the ground truth of an f32 problem is the f64 oracle on the SAME
f32-rounded inputs (`truth`).  A GPU result `got` is accepted iff

  (B) it is no farther from the truth than the f32 reference restatement is:  rel(got, truth) <= max(rel(o32, truth), FLOOR)
  (A) it agrees with the f32 reference restatement:                           rel(got, o32)   <= 1e-4
      — unless that restatement itself misses the bar on this input (rel(o32, truth) > 1e-4: an ill-conditioned
      window / group in f32), where agreeing with it to 1e-4 would mean reproducing its rounding noise; (B) alone is
      binding there, and it is the stricter statement (the GPU result must still be at least as accurate).

rel() is vector-relative: max |a - b| / max |b| (a coefficient whose true value is ~0 has no scale of its own).
FLOOR = 2e-6: two correctly rounded f32 answers can differ by a few ulps (6e-8 each, amplified by the vector norm).
"""
import numpy as np

F32_TOL = 1e-4
F64_TOL = 1e-6
FLOOR = 2e-6


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))


def accept_f32(got, o32, truth, what=""):
    eg, eo, ego = rel(got, truth), rel(o32, truth), rel(got, o32)
    assert eg <= max(eo, FLOOR), f"{what}: GPU is farther from the f64 truth ({eg:.3e}) than the f32 reference restatement ({eo:.3e})"
    if eo <= F32_TOL:
        assert ego <= F32_TOL, f"{what}: |gpu - oracle_f32| = {ego:.3e} > 1e-4 (oracle_f32 vs truth {eo:.3e})"
    return eg, eo, ego


def accept_f64(got, o64, what=""):
    e = rel(got, o64)
    assert e <= F64_TOL, f"{what}: |gpu - oracle_f64| = {e:.3e} > 1e-6"
    return e
