"""Tiny driver for ncu: runs the f32 moments kernel (K2b) a few times on a resident frame.
usage: python profiles/run_moments.py [rows] [features] [reps]"""
import sys

import torch

sys.path.insert(0, ".")
from polars_ds_extension_b200 import device as dev  # noqa: E402

rows = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
p = int(sys.argv[2]) if len(sys.argv) > 2 else 32
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
ld = (rows + 31) // 32 * 32
Z = torch.randn((p + 1, ld), device="cuda")
frame = dev.to_frame(Z, n=rows)
X, y = Z[:p], Z[p:]
M = torch.empty((p + 2, p + 2), dtype=torch.float64, device="cuda")
for _ in range(reps):
    dev.moments_frame(frame, rows, p + 1, 0, p, p, 1, out=M)
torch.cuda.synchronize()
print("moments[0,0] =", float(M[0, 0]))
