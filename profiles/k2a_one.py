"""One f64 moments shape, a few launches (ncu target): python profiles/k2a_one.py <rows> <p>"""
import sys
import torch
sys.path.insert(0, ".")
from polars_ds_extension_b200 import device as dev  # noqa: E402

n, p = int(float(sys.argv[1])), int(sys.argv[2])
Z = torch.randn((p + 1, n), device="cuda", dtype=torch.float64)
for _ in range(3):
    M = dev.moments(Z[:p], Z[p:])
torch.cuda.synchronize()
print(float(M[0, 0]))
