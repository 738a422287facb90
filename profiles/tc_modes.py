"""Accuracy + speed of the K2b variants (PDSB_TC_MODE = 0 explicit-hi, 1 raw-hi/trunc, 2 raw-hi/rna) vs float64."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from polars_ds_extension_b200 import device as dev  # noqa: E402

rows, p = 4_000_000, 32
ld = rows
g = torch.Generator(device="cuda"); g.manual_seed(1)
Z = torch.randn((p + 1, ld), device="cuda", generator=g) * 1.7 + 0.3
X, y = Z[:p], Z[p:]
M = dev.moments(X, y, n=rows).cpu().numpy()
Zd = torch.cat([Z.double(), torch.ones((1, ld), dtype=torch.float64, device="cuda")])
ref = (Zd @ Zd.T).cpu().numpy()
sc = np.sqrt(np.outer(np.diag(ref), np.diag(ref)))
err = np.nan_to_num(np.abs(M - ref) / sc)
big = 50_000_000 // 128 * 128
Zb = torch.randn((p + 1, big), device="cuda")
Mb = torch.empty((p + 2, p + 2), dtype=torch.float64, device="cuda")
use_frame = os.environ.get("PDSB_FRAME", "1") != "0"
if use_frame:
    Fb = dev.to_frame(Zb, n=big)
    run = lambda: dev.moments_frame(Fb, big, p + 1, 0, p, p, 1, out=Mb)
else:
    run = lambda: dev.moments(Zb[:p], Zb[p:], n=big, out=Mb)
for _ in range(3):
    run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"variant={os.environ.get('PDSB_TC_MODE','1')} ring={os.environ.get('PDSB_TC_RING','0')} dbg={os.environ.get('PDSB_TC_DBG','0')} frame={int(use_frame)} max_rel_err={err.max():.3e} "
      f"ms(5e7 rows)={ms:.3f} GB/s={big*(p+1)*4/ms/1e6:.0f}")
