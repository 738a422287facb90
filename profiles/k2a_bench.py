"""K2a (generic SIMT moments) timings: f64 frames, f32 forced onto the SIMT path, weighted, and a small-n latency."""
import sys
import torch
sys.path.insert(0, ".")
from polars_ds_extension_b200 import device as dev  # noqa: E402
from polars_ds_extension_b200._lib import lib  # noqa: E402

PEAK = 6570.0


def ev(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


import os
SHAPES = [(torch.float64, 20_000_000, 32), (torch.float32, 40_000_000, 32), (torch.float64, 50_000_000, 8),
          (torch.float64, 100_000, 4)]
if os.environ.get("K2A_F64_ONLY"):
    SHAPES = [(torch.float64, 20_000_000, 32), (torch.float32, 40_000_000, 32), (torch.float64, 20_000_000, 30), (torch.float64, 50_000_000, 8), (torch.float64, 30_000_000, 16),
              (torch.float64, 10_000_000, 62), (torch.float64, 10_000_000, 64), (torch.float64, 100_000, 4)]
for dtype, n, p in SHAPES:
    Z = torch.randn((p + 1, n), device="cuda", dtype=dtype)
    w = torch.rand(n, device="cuda", dtype=dtype) + 0.5
    lib().pdsb_set_moments_path(1)          # force K2a also for f32
    ms = ev(lambda: dev.moments(Z[:p], Z[p:]))
    msw = ev(lambda: dev.moments(Z[:p], Z[p:], w=w))
    lib().pdsb_set_moments_path(0)
    gb = n * (p + 1) * Z.element_size() / 1e6
    ref = (Z.double() @ Z.double().T) if n <= 20_000_003 else None
    err = float(((dev.moments(Z[:p], Z[p:])[: p + 1, : p + 1] - ref).abs() / ref.abs().clamp_min(1.0)).max()) if ref is not None else -1
    print(f"K2a {str(dtype)[6:]} n={n:.0e} p={p}: {ms:.3f} ms -> {gb / ms:.0f} GB/s ({gb / ms / PEAK * 100:.1f}% of HBM); weighted {msw:.3f} ms; max rel err {err:.2e}")
    del Z, w
