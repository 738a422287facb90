"""Worker of tests/test_gpu_multi.py::test_world_communicator_fits_one_regression_over_all_ranks (run under torchrun)."""
import os
import sys

import numpy as np
import pyarrow as pa
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from polars_ds_extension_b200 import _harness, parallel  # noqa: E402
from polars_ds_extension_b200._lib import check, lib  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
check(lib().pdsb_set_device(local))
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
assert parallel.init_world() == world and lib().pdsb_comm_size() == world

n, p = 400_000 + 1000 * rank, 6
rng = np.random.default_rng(100)                      # every rank draws the SAME global data, then keeps its rows
sizes = [400_000 + 1000 * r for r in range(world)]
X = rng.standard_normal((p, sum(sizes))).astype(np.float32)
beta = ((np.arange(p) % 7) - 3) / 4.0
y = (beta @ X + 0.25 + 0.1 * rng.standard_normal(sum(sizes))).astype(np.float32)
lo = sum(sizes[:rank])
Xr, yr = X[:, lo:lo + n], y[lo:lo + n]
kw = {"bias": True, "null_policy": "skip", "l1_reg": 0.0, "l2_reg": 0.0, "solver": "qr", "tol": 1e-5, "max_iter": 200,
      "weighted": False, "positive": False, "singular_x_tol": 1e-6}
names = ["y"] + [f"x{i}" for i in range(p)]
cols = [pa.array(yr)] + [pa.array(np.ascontiguousarray(Xr[i])) for i in range(p)]
c = np.asarray(_harness.call_plugin("pl_lr_f32", cols, names, kw)[0].as_py())
pr = _harness.call_plugin("pl_lr_pred_f32", cols, names, kw).field("pred").to_numpy(zero_copy_only=False)
A = np.column_stack([X.T.astype(np.float64), np.ones(X.shape[1])])
ref, *_ = np.linalg.lstsq(A, y.astype(np.float64), rcond=None)
coef_err = np.max(np.abs(c - ref)) / np.max(np.abs(ref))
pred_ref = A[lo:lo + n] @ ref
pred_err = np.max(np.abs(pr - pred_ref)) / np.max(np.abs(pred_ref))
t = torch.tensor([coef_err, pred_err, 1.0 if len(pr) == n else 0.0], dtype=torch.float64, device="cuda")
gathered = [torch.empty_like(t) for _ in range(world)]
dist.all_gather(gathered, t)
parallel.destroy_world()
dist.destroy_process_group()
if rank == 0:
    g = torch.stack(gathered).cpu().numpy()
    np.savez(sys.argv[1], coef_err=g[:, 0], pred_err=g[:, 1], ok=g[:, 2] > 0.5)
