"""One plugin call, k GPUs (single process): e2e rows/s of `_polars_plugin_pl_lr_pred_f32` on 1e8 x 32 f32 host buffers
when the library row-shards the call over a device group (pdsb_set_devices, include/pdsb.h) — every shard goes up its
own PCIe link, the f64 moments are all-reduced by the in-process NCCL communicator, each GPU predicts its rows.

    python profiles/e2e_devices.py [rows] [features] > profiles/e2e_devices_r02.jsonl      (gpurun --gpus 8)
"""
import json
import sys
import time

import numpy as np
import pyarrow as pa
import torch

sys.path.insert(0, ".")
from polars_ds_extension_b200 import _harness, parallel  # noqa: E402
from polars_ds_extension_b200._lib import lib  # noqa: E402

rows = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
p = int(sys.argv[2]) if len(sys.argv) > 2 else 32
KW = {"bias": False, "null_policy": "skip", "l1_reg": 0.0, "l2_reg": 0.0, "solver": "qr", "tol": 1e-5, "max_iter": 200,
      "weighted": False, "positive": False, "singular_x_tol": 1e-6}

sys.path.insert(0, ".")
from bench import gen_host_lin_reg  # noqa: E402

cols = gen_host_lin_reg(rows, p)
names = ["y"] + [f"x{i}" for i in range(p)]
pinned = None
ngpu = torch.cuda.device_count()
base = None
for k in [1, 2, 4, 8]:
    if k > ngpu:
        break
    parallel.set_devices(list(range(k)) if k > 1 else [])
    for kind in ("pageable", "pinned"):
        if kind == "pinned" and pinned is None:
            pinned = []
            for c in cols:
                t = torch.empty(c.shape, dtype=torch.float32, pin_memory=True)
                t.numpy()[:] = c
                pinned.append(t.numpy())
        src = pinned if kind == "pinned" else cols
        inputs = [pa.array(a) for a in src]
        for _ in range(2):
            r = _harness.call_plugin("pl_lr_pred_f32", inputs, names, KW)
        pred0 = r.field("pred").to_numpy(zero_copy_only=False)[:1000].copy()
        del r
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            r = _harness.call_plugin("pl_lr_pred_f32", inputs, names, KW)
            assert len(r) == rows
            del r
        dt = (time.perf_counter() - t0) / reps
        if base is None:
            base = pred0
        print(json.dumps({"devices": k, "inputs": kind, "rows": rows, "features": p, "ms_per_call": dt * 1e3,
                          "rows_per_s": rows / dt, "h2d_GBps": (p + 1) * rows * 4 / dt / 1e9,
                          "staged_bytes": int(lib().pdsb_last_staged_bytes()),
                          "pred_max_abs_diff_vs_1gpu_first_1000": float(np.max(np.abs(pred0 - base)))}), flush=True)
parallel.set_devices([])
