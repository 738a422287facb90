"""Ring-depth sweep of the f64 side / wide moments kernels (sweep build of the library, PDSB_K2A_DEPTH read per call):
   K2B_LIB=profiles/_ab/lib_k2a_sweep.so python profiles/k2a_depth_sweep.py
One line per (p, targets, weighted, depth): ms and GB/s of the moments call (Gram + partial reduce)."""
import os
import sys
from pathlib import Path
import torch
sys.path.insert(0, ".")
if os.environ.get("K2B_LIB"):
    from polars_ds_extension_b200 import _lib as _libmod
    _libmod.LIB_PATH = Path(os.environ["K2B_LIB"]).resolve()
from polars_ds_extension_b200 import device as dev  # noqa: E402


def ev(fn, reps=4):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# (p, t): t = 1 and p % 8 in {0, 7} -> side kernel with ceil(p / 8) blocks; otherwise wide with ceil((p + t + 1) / 8)
SHAPES = [(8, 1), (16, 1), (24, 1), (32, 1), (40, 1), (48, 1), (56, 1), (64, 1),
          (6, 1), (14, 1), (22, 1), (30, 1), (38, 1), (46, 1), (54, 1), (62, 1)]
for p, t in SHAPES:
    n = int(4e9 / (8 * (p + t))) // 64 * 64
    Z = torch.randn((p + t, n), device="cuda", dtype=torch.float64)
    w = torch.rand(n, device="cuda", dtype=torch.float64) + 0.5
    for weighted in (False, True):
        best = None
        for depth in [0, 2, 3, 4, 5, 6, 7, 8, 9]:
            os.environ["PDSB_K2A_DEPTH"] = str(depth)
            ms = ev(lambda: dev.moments(Z[:p], Z[p:], w=w if weighted else None))
            gbs = n * (p + t + (1 if weighted else 0)) * 8 / ms / 1e6
            print(f"p={p} t={t} weighted={int(weighted)} depth={depth}: {ms:.3f} ms {gbs:.0f} GB/s", flush=True)
            if depth and (best is None or ms < best[1]):
                best = (depth, ms, gbs)
        print(f"BEST p={p} t={t} weighted={int(weighted)} depth={best[0]} {best[2]:.0f} GB/s ({best[2] / 65.7:.1f} %)", flush=True)
    del Z, w
