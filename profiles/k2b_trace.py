"""Timeline of CTA 0 of the K2b kernel (ablation build with PDSB_TC_DBG=16): clock64 stamps per stage and role, reduced to
the steady-state intervals that say where a stage's latency goes.
usage: K2B_LIB=profiles/_ab/lib_trace.so PDSB_TC_DBG=16 python profiles/k2b_trace.py [rows] [features]"""
import ctypes as C
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, ".")
import polars_ds_extension_b200._lib as _libmod  # noqa: E402

_libmod.LIB_PATH = Path(os.environ.get("K2B_LIB", "profiles/_ab/lib_trace.so")).resolve()
from polars_ds_extension_b200 import device as dev  # noqa: E402
from polars_ds_extension_b200._lib import lib  # noqa: E402

rows = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
p = int(sys.argv[2]) if len(sys.argv) > 2 else 32
ld = (rows + 31) // 32 * 32
Z = torch.randn((p + 1, ld), device="cuda")
frame = dev.to_frame(Z, n=rows)
del Z
M = torch.empty((p + 2, p + 2), dtype=torch.float64, device="cuda")
for _ in range(3):
    dev.moments_frame(frame, rows, p + 1, 0, p, p, 1, out=M)
torch.cuda.synchronize()
NS = 2048
buf = np.zeros((NS, 16), dtype=np.uint64)
L = lib()
L.pdsb_debug_tc_trace.argtypes = [C.c_void_p, C.c_int]
L.pdsb_debug_tc_trace.restype = C.c_int
ne = L.pdsb_debug_tc_trace(buf.ctypes.data, NS)
assert ne == 16, ne
t = buf.astype(np.int64)
lo, hi = 200, 1800          # steady state


def d(a, b):
    x = (t[lo:hi, b] - t[lo:hi, a]).astype(np.float64)
    return {"mean": round(float(np.mean(x)), 1), "p10": float(np.percentile(x, 10)), "p90": float(np.percentile(x, 90))}


period = float(np.mean(np.diff(t[lo:hi, 6])))
ring = {1: 10, 2: 10, 3: 7, 4: 5}[(p + 15) // 16]
out = {"rows": rows, "p": p, "cycles_per_stage": round(period, 1),
       "tma_issue->conv_sees_tile (HBM latency + queueing)": d(0, 2),
       "conv_sees_tile->conv_has_A_slot": d(2, 3),
       "conv work (LDS, lo, STTM issue)": d(3, 4),
       "conv wait::st + fences + arrive": d(4, 5),
       "conv total": d(2, 5),
       "a_full arrive->mma wakes": d(5, 6),
       "mma issue + commits": d(6, 7),
       "mma wakes->slot free seen by producer (stage it+RING)": round(float(np.mean(t[lo + ring:hi + ring, 0] - t[lo:hi, 6])), 1),
       "slot hold: tma issue(it) -> slot free (it+RING)": round(float(np.mean(t[lo + ring:hi + ring, 0] - t[lo:hi, 0])), 1),
       "producer stall: slot free(it) - slot free(it-1)": round(float(np.mean(np.diff(t[lo:hi, 0]))), 1),
       }
out["side warp 0: loop top -> tile seen"] = d(15, 12)
out["side warp 0: work (2 boxes) + arrive"] = d(12, 13)
out["side warp 0: period"] = round(float(np.mean(np.diff(t[lo:hi, 12]))), 1)
out["side warp 0 lag behind the converter (tile seen)"] = d(2, 12)
out["producer: slot free -> TMA issued"] = d(0, 14)
e = t[lo:hi, :]
e = e[(e[:, 8] > 0) & (e[:, 9] > 0)]
if len(e):
    out["epi: d_full seen -> d_empty arrived (tcgen05.ld)"] = round(float(np.mean(e[:, 9] - e[:, 8])), 1)
    out["epi: f64 adds"] = round(float(np.mean(e[:, 10] - e[:, 9])), 1)
    out["mma commit -> epi sees d_full"] = round(float(np.mean(e[:, 8] - e[:, 7])), 1)
print(json.dumps(out, indent=1))
np.save("gpurun_out/k2b_trace_p%d.npy" % p, t)
