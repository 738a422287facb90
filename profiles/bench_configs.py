"""Device-resident timings of the other BASELINE.json configs (parity-test cases, not the bench line):
C1 lin_reg 100k x 4 f64 (through the plugin ABI, latency), C3 grouped 1e4 groups x ~1e4 rows x 8 f32,
C4 rolling window 1024 on 1e8 x 8 f32, C4' recursive on the same frame.  Prints one JSON line per config."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import polars_ds_extension_b200 as pds  # noqa: E402
from polars_ds_extension_b200 import device as dev  # noqa: E402

PEAK = 6570.0


def ev_time(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
out = []

# ---- C1: through the plugin ABI (host buffers), f64 ----
rng = np.random.default_rng(208)
n, p = 100_000, 4
X = rng.standard_normal((n, p))
y = X @ [0.5, -0.25, 0.75, 0.1] + 0.5 + 0.1 * rng.standard_normal(n)
df = pds.Frame({f"x{i}": X[:, i] for i in range(p)} | {"y": y})
e = pds.lin_reg(*[f"x{i}" for i in range(p)], target="y", add_bias=True)
for _ in range(3):
    df.select(e)
t0 = time.perf_counter()
for _ in range(20):
    r = df.select(e)
dt = (time.perf_counter() - t0) / 20
ref, *_ = np.linalg.lstsq(np.column_stack([X, np.ones(n)]), y, rcond=None)
err = float(np.max(np.abs(np.asarray(r["coeffs"][0].as_py()) - ref)))
out.append({"config": "C1 pds.lin_reg 100k x 4 f64 add_bias (plugin ABI, host buffers)", "ms": dt * 1e3, "rows_per_s": n / dt,
            "max_abs_err_vs_lstsq": err})

# ---- C3: grouped ----
n = int(1e8 * scale)
p = 8
g = torch.Generator(device="cuda"); g.manual_seed(208)
Z = torch.randn((p + 1, n), device="cuda", generator=g)
sizes = torch.randint(8000, 12001, (int(n / 10000) + 2,), generator=torch.Generator().manual_seed(1))
offs = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(sizes, 0)])
offs = offs[offs < n]
offs = torch.cat([offs, torch.tensor([n])]).cuda()
ms = ev_time(lambda: dev.grouped_lin_reg(Z[:p], Z[p], offs, add_bias=True, singular_x_tol=1e-6))
bytes_ = n * (p + 1) * 4
out.append({"config": f"C3 grouped lin_reg {offs.numel() - 1} groups x ~1e4 rows x 8 f32 (+bias), device-resident", "ms": ms,
            "rows_per_s": n / (ms * 1e-3), "GBps": bytes_ / ms / 1e6, "frac_hbm": bytes_ / ms / 1e6 / PEAK})

# ---- C4: rolling / recursive ----
coeffs = torch.empty((n, p), dtype=torch.float32, device="cuda")
pred = torch.empty(n, dtype=torch.float32, device="cuda")
valid = torch.empty(n, dtype=torch.uint8, device="cuda")
ms = ev_time(lambda: dev.online_lin_reg(Z[:p], Z[p], 1024, p, coeffs=coeffs, pred=pred, valid=valid), reps=3, warm=1)
bytes_ = n * ((p + 1) * 4 + p * 4 + 4 + 1)
out.append({"config": "C4 rolling_lin_reg window=1024 on 1e8 x 8 f32, device-resident", "ms": ms, "rows_per_s": n / (ms * 1e-3),
            "GBps": bytes_ / ms / 1e6, "frac_hbm": bytes_ / ms / 1e6 / PEAK})
# spot check against the definition
j = 123_456
Xw = Z[:p, j - 1023:j + 1].double().T.cpu().numpy(); yw = Z[p, j - 1023:j + 1].double().cpu().numpy()
ref = np.linalg.solve(Xw.T @ Xw, Xw.T @ yw)
out[-1]["max_abs_err_row_123456"] = float(np.max(np.abs(coeffs[j].cpu().numpy() - ref)))
ms = ev_time(lambda: dev.online_lin_reg(Z[:p], Z[p], 0, p, coeffs=coeffs, pred=pred, valid=valid), reps=3, warm=1)
out.append({"config": "C4' recursive_lin_reg on 1e8 x 8 f32, device-resident", "ms": ms, "rows_per_s": n / (ms * 1e-3),
            "GBps": bytes_ / ms / 1e6, "frac_hbm": bytes_ / ms / 1e6 / PEAK})
for o in out:
    print(json.dumps(o))
