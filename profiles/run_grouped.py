"""ncu target: grouped lin_reg (K5) on n x 8 f32 (+bias), ~1e4-row groups, device-resident."""
import sys
import torch
sys.path.insert(0, ".")
from polars_ds_extension_b200 import device as dev  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
p = 8
g = torch.Generator(device="cuda"); g.manual_seed(208)
Z = torch.randn((p + 1, n), device="cuda", generator=g)
sizes = torch.randint(8000, 12001, (int(n / 10000) + 2,), generator=torch.Generator().manual_seed(1))
offs = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(sizes, 0)])
offs = offs[offs < n]
offs = torch.cat([offs, torch.tensor([n])]).cuda()
for _ in range(3):
    dev.grouped_lin_reg(Z[:p], Z[p], offs, add_bias=True)
torch.cuda.synchronize()
