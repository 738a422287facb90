"""One f64 moments shape, a few launches (ncu target): python profiles/k2a_one.py <rows> <p>"""
import os
import sys
from pathlib import Path
import torch
sys.path.insert(0, ".")
if os.environ.get("K2B_LIB"):          # A/B against a kept build of the library (profiles/_ab/*.so, not in git)
    from polars_ds_extension_b200 import _lib as _libmod
    _libmod.LIB_PATH = Path(os.environ["K2B_LIB"]).resolve()
from polars_ds_extension_b200 import device as dev  # noqa: E402

n, p = int(float(sys.argv[1])), int(sys.argv[2])
Z = torch.randn((p + 1, n), device="cuda", dtype=torch.float64)
for _ in range(3):
    M = dev.moments(Z[:p], Z[p:])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    M = dev.moments(Z[:p], Z[p:])
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"f64 n={n:.0e} p={p} cfg={os.environ.get('PDSB_K2A_CFG', '0')}: {ms:.3f} ms, {n * (p + 1) * 8 / ms / 1e6:.0f} GB/s, M00={float(M[0, 0]):.6e}")
