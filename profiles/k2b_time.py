"""CUDA-event timing of the f32 moments kernel (K2b) on a resident frame and on the column-major matrix, plus the
column-major -> frame conversion.  Kernel knobs are environment variables read once per process (PDSB_TC_*), so a
sweep runs this script once per setting.
usage: python profiles/k2b_time.py [rows] [features] [reps]"""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
import polars_ds_extension_b200._lib as _libmod  # noqa: E402

if os.environ.get("K2B_LIB"):          # A/B against a kept build of the library (profiles/_ab/*.so, not in git)
    from pathlib import Path

    _libmod.LIB_PATH = Path(os.environ["K2B_LIB"]).resolve()
from polars_ds_extension_b200 import device as dev  # noqa: E402
from polars_ds_extension_b200._lib import lib  # noqa: E402

rows = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
p = int(sys.argv[2]) if len(sys.argv) > 2 else 32
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
ld = (rows + 31) // 32 * 32
Z = torch.randn((p + 1, ld), device="cuda")
frame = dev.to_frame(Z, n=rows)
X, y = Z[:p], Z[p:]
M = torch.empty((p + 2, p + 2), dtype=torch.float64, device="cuda")
M2 = torch.empty_like(M)


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


t_frame = timed(lambda: dev.moments_frame(frame, rows, p + 1, 0, p, p, 1, out=M))
t_col = timed(lambda: dev.moments(X, y, n=rows, out=M2))
t_conv = timed(lambda: dev.to_frame(Z, n=rows, out=frame)) if "out" in dev.to_frame.__code__.co_varnames else timed(lambda: dev.to_frame(Z, n=rows))
gb = rows * (p + 1) * 4 / 1e9
# independent check: exact-product f64-accumulated SIMT moments of a 2e6-row prefix vs the tensor-core kernel on the same prefix
m = min(rows, 2_000_000)
a = dev.moments(X, y, n=m).cpu()
lib().pdsb_set_moments_path(1)
b = dev.moments(X, y, n=m).cpu()
lib().pdsb_set_moments_path(0)
rel = float(((a - b).abs() / b.abs().clamp_min(1e-300)).max())
same = bool(torch.equal(M.cpu(), M2.cpu()))
print(json.dumps({"rows": rows, "p": p, "env": {k: v for k, v in os.environ.items() if k.startswith(("PDSB_", "K2B_"))},
                  "frame_ms": t_frame, "frame_GBs": gb / t_frame * 1e3, "colmajor_ms": t_col, "colmajor_GBs": gb / t_col * 1e3,
                  "to_frame_ms": t_conv, "to_frame_GBs": 2 * gb / t_conv * 1e3, "frame_equals_colmajor_bits": same,
                  "max_rel_vs_simt_f64acc_2e6_rows": rel}))
