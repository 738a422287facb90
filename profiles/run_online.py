"""ncu target: one rolling_lin_reg (window 1024) and one recursive_lin_reg on n x 8 f32 (device-resident)."""
import sys
import torch
sys.path.insert(0, ".")
from polars_ds_extension_b200 import device as dev  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
g = torch.Generator(device="cuda"); g.manual_seed(208)
Z = torch.randn((9, n), device="cuda", generator=g)
for _ in range(2):
    dev.online_lin_reg(Z[:8], Z[8], window=1024, min_rows=1024, add_bias=True)
    dev.online_lin_reg(Z[:8], Z[8], window=0, min_rows=9, add_bias=True)
torch.cuda.synchronize()
