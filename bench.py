#!/usr/bin/env python
"""bench.py — BASELINE.json's metric: pds.lin_reg rows/sec on 1e8 x 32 f32, return_pred=True (configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--rows R] [--features P] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic rows: moments (K2) -> solve (K3) ->
predict/residual (K4).  Prints ONE JSON line (rank 0):

  value      whole-job rows/s with the frame already resident in HBM (CUDA events, max over ranks)
  e2e        the same metric through the reference-facing plugin symbol `_polars_plugin_pl_lr_pred_f32` with HOST
             (pinned) Arrow buffers: H2D of the 33 columns and D2H of pred+resid are inside the timed region
  roofline   the dominant kernel (the Gram/moments kernel) timed alone with CUDA events; algorithmic bytes =
             (p + 1) * 4 per row against the measured HBM peak in MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle ("port" of the reference, numpy/OpenBLAS with all host threads) on a bounded sample

N > 1 (torchrun): rows are sharded, every rank owns `--rows` rows (weak scaling); per step each rank builds its
partial moments, ONE NCCL all-reduce sums the (p+2)^2 f64 moments, every rank solves redundantly and predicts its
shard.  `--impl reference` times the oracle port on the host cores (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--rows", type=int, default=100_000_000, help="rows per GPU")
    ap.add_argument("--features", type=int, default=32)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-rows", type=int, default=8_000_000, help="bounded CPU-baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi sampler running during the timed region (B200_PROFILING.md clocks line)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, uuid: str | None = None):
        self.index = index
        self.uuid = uuid          # "GPU-..." of the CUDA device in use: immune to CUDA_VISIBLE_DEVICES re-numbering
        self.samples = []
        self.proc = None
        self.nvml = []            # (sm_mhz, reason bitmask) every ~2 ms from NVML, when the library is loadable
        self.nvml_max = None
        self._run = False

    def _nvml_loop(self):
        # in-process NVML polling: nvidia-smi's own loop cannot go below ~100 ms, shorter than one default bench run
        try:
            import pynvml

            pynvml.nvmlInit()
            h = None
            if self.uuid:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(self.uuid.encode() if isinstance(self.uuid, str) else self.uuid)
                except Exception:
                    h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.nvml_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            while self._run:
                mhz = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                try:
                    bits = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    bits = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.nvml.append((mhz, bits))
                time.sleep(0.002)
        except Exception:
            pass

    def start(self):
        try:
            self._run = True
            self.nvml_thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.nvml_thread.start()
        except Exception:
            self._run = False
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        self._run = False
        try:
            out = self._stop_smi()
        except Exception as e:      # never let the sampler take the bench line down
            out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"sampler error: {e}"], "samples": 0}
        try:
            nv = list(self.nvml)
            smi_mhz = out.get("sm_mhz")
            nv_mhz = float(np.median([m for m, _ in nv])) if nv else None
            if len(nv) >= 3 and smi_mhz and abs(nv_mhz - smi_mhz) > 0.25 * smi_mhz:
                out["note"] = f"NVML samples ({nv_mhz:.0f} MHz) disagree with nvidia-smi; nvidia-smi reported"
            elif len(nv) >= 3:
                bits = 0
                for _, b in nv:
                    bits |= b
                names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
                reasons = set(out.get("reasons") or []) | {nm for m, nm in names.items() if bits & m}
                reasons.discard("nvidia-smi unavailable")
                out = {"sm_mhz": float(np.median([m for m, _ in nv])), "sm_max_mhz": self.nvml_max or out.get("sm_max_mhz"),
                       "reasons": sorted(reasons), "samples": len(nv), "source": "nvml (2 ms period) + nvidia-smi"}
        except Exception:
            pass
        return out

    def _stop_smi(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            parts = [x.strip() for x in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def synth_on_device(torch, rows, p, seed, device):
    """X ~ N(0,1), beta_j = ((j mod 7) - 3)/4, y = X beta + 0.1 N(0,1)  (seed 208 as the reference's
    benchmarks/test_linear_regression.py:9).  Column-major: tensor (p, ld)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    ld = (rows + 31) // 32 * 32
    # one column-major frame [X | y] (the layout the plugin's packer builds; lets the TMA box cover all q columns)
    Z = torch.zeros((p + 1, ld), dtype=torch.float32, device=device)
    X, y = Z[:p], Z[p:]
    beta = ((torch.arange(p, device=device) % 7).float() - 3.0) / 4.0
    chunk = 1 << 24
    for c in range(p):
        X[c].normal_(generator=g)
    for s in range(0, ld, chunk):
        e = min(ld, s + chunk)
        y[0, s:e] = beta @ X[:, s:e]
    noise = torch.empty(ld, dtype=torch.float32, device=device).normal_(generator=g)
    y[0] += 0.1 * noise
    del noise
    return X, y, ld


def cpu_baseline(rows, p, steps=None, warmup=1, min_seconds=10.0):
    """The oracle port (numpy/OpenBLAS, all host threads) on a bounded sample of the same workload.
    steps=None: repeat the sample until `min_seconds` of CPU work have been timed (at least 3 passes)."""
    from oracle import lin_reg_oracle as orc

    rng = np.random.default_rng(208)
    X = rng.standard_normal((p, rows), dtype=np.float32)
    beta = ((np.arange(p) % 7) - 3.0) / 4.0
    y = (beta.astype(np.float32) @ X + 0.1 * rng.standard_normal(rows, dtype=np.float32)).astype(np.float32)
    cols = [orc.Col("y", y)] + [orc.Col(f"x{i}", X[i]) for i in range(p)]
    kw = {"bias": False, "null_policy": "skip", "l1_reg": 0.0, "l2_reg": 0.0, "solver": "qr", "tol": 1e-5,
          "max_iter": 200, "weighted": False, "positive": False, "singular_x_tol": 1e-6}
    for _ in range(max(1, warmup)):
        orc.pl_lr_pred(cols[:], kw, f32=True)  # warm-up (BLAS threads, page faults)
    t0 = time.perf_counter()
    done = 0
    while True:
        out = orc.pl_lr_pred(cols, kw, f32=True)
        done += 1
        if steps is not None and done >= steps:
            break
        if steps is None and done >= 3 and time.perf_counter() - t0 >= min_seconds:
            break
    dt = (time.perf_counter() - t0) / done
    assert out["pred"][0].shape[0] == rows
    return rows / dt, dt, done


def run_reference(args, rank):
    if rank != 0:
        return
    rows = args.cpu_rows
    v, dt, done = cpu_baseline(rows, args.features, steps=args.steps, warmup=args.warmup)
    cores = os.cpu_count()
    line = {
        "impl": "reference", "metric": "lin_reg rows/sec (f32, return_pred=True)", "value": v, "unit": "rows/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"pds.lin_reg {args.rows} rows x {args.features} f32 features per GPU, add_bias=False, "
                               f"return_pred=True (BASELINE configs[1]); step = moments + solve + predict/resid",
                   "rows_per_gpu": args.rows, "features": args.features,
                   "sample": f"each CPU step = the same expression on a {rows}-row sample of that frame"},
        "cpu_baseline": {"value": v, "unit": "rows/s", "cores": cores, "kind": "port",
                         "sample": f"{done} passes over {rows} rows x {args.features} f32 through oracle.pl_lr_pred "
                                   f"(numpy/OpenBLAS, all host threads), {dt:.2f} s each"},
        "e2e": {"value": v, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    from polars_ds_extension_b200 import device as dev
    from polars_ds_extension_b200._lib import lib, METHOD_LSTSQ

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    from polars_ds_extension_b200._lib import check
    check(lib().pdsb_set_device(local_rank))
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    n_gpus = world
    rows, p = args.rows, args.features
    L = lib()

    X, y, ld = synth_on_device(torch, rows, p, 208 + rank, device)
    e2e_host = None
    if not args.no_e2e:                                            # every rank pushes its own shard through the plugin
        e2e_host = stage_host_copy(torch, X, y, rows, p)          # column-major host (Arrow) buffers for the e2e leg
    # resident layout of the hot path: the library's row-blocked frame ([block][column][128], include/pdsb.h)
    Zcm = X._base if X._base is not None else torch.cat([X, y])
    frame = dev.to_frame(Zcm, n=rows)
    torch.cuda.synchronize()
    del X, y, Zcm
    torch.cuda.empty_cache()
    ncols = p + 1
    q1 = p + 2
    M = torch.empty((q1, q1), dtype=torch.float64, device=device)
    beta = torch.empty((1, p), dtype=torch.float64, device=device)
    status = torch.zeros(4, dtype=torch.int32, device=device)
    pred = torch.empty((1, ld), dtype=torch.float32, device=device)
    resid = torch.empty((1, ld), dtype=torch.float32, device=device)
    tol = 1e-6  # default singular_x_tol of the f32 family (expr_linear.py:184-186)

    def step():
        dev.moments_frame(frame, rows, ncols, 0, p, p, 1, out=M)
        if world > 1:
            dist.all_reduce(M)  # the only exchange of the row-sharded path: (p+2)^2 f64 partial moments
        dev.solve(M, p, 1, add_bias=False, method=METHOD_LSTSQ, singular_x_tol=tol, beta=beta, status=status)
        dev.predict_frame(frame, rows, ncols, 0, p, p, 1, beta, status, False, pred, resid)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    gpu_uuid = None
    try:
        u = str(torch.cuda.get_device_properties(device).uuid)
        gpu_uuid = u if u.startswith("GPU-") else "GPU-" + u
    except Exception:
        gpu_uuid = None
    sampler = ClockSampler(local_rank, gpu_uuid)
    if rank == 0:
        sampler.start()
    launches0 = dev.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    launches = dev.launch_count() - launches0
    ms = ev0.elapsed_time(ev1)
    t_ms = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    clocks = sampler.stop() if rank == 0 else None
    value = rows * n_gpus * args.steps / (ms * 1e-3)
    path = int(L.pdsb_last_moments_path())

    # ---- roofline of the dominant kernel (moments), timed alone ----
    for _ in range(3):
        dev.moments_frame(frame, rows, ncols, 0, p, p, 1, out=M)
    torch.cuda.synchronize()
    reps = max(args.steps, 5)
    ev0.record()
    for _ in range(reps):
        dev.moments_frame(frame, rows, ncols, 0, p, p, 1, out=M)
    ev1.record()
    torch.cuda.synchronize()
    k_ms = ev0.elapsed_time(ev1) / reps
    alg_bytes = rows * (p + 1) * 4
    # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this kernel on this workload, from the committed
    # `ncu --set full` capture (profiles/gram_tcgen05_r01_ncu_metrics.csv: 13.200198 GB + 3.690496 MB); only valid for
    # the default shape, null otherwise
    traffic = 13_203_888_496 if (rows == 100_000_000 and p == 32 and path == 1) else None
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f)["hbm_gbs"])
            peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        pass
    # coefficients sanity inside the bench: parity against the generating beta (noise 0.1 => tiny error at 1e8 rows)
    bt = ((np.arange(p) % 7) - 3.0) / 4.0
    coef_err = float(np.max(np.abs(beta.cpu().numpy()[0] - bt)))

    e2e = None
    if not args.no_e2e:
        def all_ok(flag):
            torch.cuda.synchronize()
            if world == 1:
                return bool(flag)
            t = torch.tensor([1.0 if flag else 0.0], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            return float(t.item()) > 0.5

        e2e = run_e2e(torch, e2e_host, rows, p, args, all_ok)
        if world > 1:                                   # unconditional on every rank (value None -> contributes 0)
            t_e = torch.tensor([e2e["ms_per_step"] if e2e.get("value") else 0.0], dtype=torch.float64, device=device)
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
            if e2e.get("value"):
                e2e["ms_per_step"] = float(t_e.item())
                e2e["value"] = rows * n_gpus / (e2e["ms_per_step"] * 1e-3)
                e2e["h2d_bytes_per_step"] *= n_gpus
                e2e["d2h_bytes_per_step"] *= n_gpus
                e2e["api"] += f"; {n_gpus} ranks concurrently, max over ranks"

    cpu = None
    if rank == 0 and not args.no_cpu and n_gpus == 1:
        v, dt, done = cpu_baseline(args.cpu_rows, p)
        cpu = {"value": v, "unit": "rows/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"{done} passes over {args.cpu_rows} rows x {p} f32 through oracle.pl_lr_pred (numpy/OpenBLAS, "
                         f"all host threads), {dt:.2f} s each"}

    if rank == 0:
        line = {
            "metric": "lin_reg rows/sec (f32, return_pred=True)", "value": value, "unit": "rows/s", "n_gpus": n_gpus,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"pds.lin_reg {rows} rows x {p} f32 features per GPU, add_bias=False, return_pred=True "
                                   f"(BASELINE configs[1]); step = moments + solve + predict/resid",
                       "rows_per_gpu": rows, "features": p,
                       "parallelism": f"row-sharded x{n_gpus}, one f64 moments all-reduce per step" if n_gpus > 1 else "single GPU",
                       "l2_policy": "inputs (13.2 GB per step) are larger than L2; no explicit flush",
                       "resident_layout": "row-blocked frame [block][column][128] (library native, include/pdsb.h)",
                       "moments_kernel": "tcgen05+TMA 3xTF32" if path == 1 else "simt f32 (f64 accumulate)",
                       "max_abs_coef_error_vs_generating_beta": coef_err},
            "e2e": e2e,
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "moments (Gram X'X | X'y)", "kernel_ms": k_ms,
                         "algorithmic_bytes": alg_bytes, "peak_source": peak_src},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def stage_host_copy(torch, X, y, rows, p):
    """Pinned host copy of the synthetic frame as p+1 separate column buffers (what Arrow hands the plugin)."""
    try:
        host = torch.empty((p + 1, rows), dtype=torch.float32, pin_memory=True)
    except Exception as e:  # not enough lockable host memory
        return f"pinned allocation failed: {e}"
    host[0].copy_(y[0, :rows])
    for c in range(p):
        host[c + 1].copy_(X[c, :rows])
    torch.cuda.synchronize()
    return host


def run_e2e(torch, host, rows, p, args, all_ok=None):
    """Through the plugin C ABI with host buffers: what a Polars user of the drop-in library would time.
    `all_ok(flag) -> bool` is a collective AND over the ranks (and a barrier); every rank calls it exactly twice,
    whatever happens locally, so a failing rank can never leave the others waiting."""
    import pyarrow as pa

    from polars_ds_extension_b200 import _harness

    all_ok = all_ok or (lambda flag: bool(flag))
    err = host if isinstance(host, str) else None
    inputs = names = kw = None
    if err is None:
        try:
            hn = host.numpy()
            inputs = [pa.array(hn[i]) for i in range(p + 1)]          # zero-copy views of the pinned buffers
            names = ["y"] + [f"x{i}" for i in range(p)]
            kw = {"bias": False, "null_policy": "skip", "l1_reg": 0.0, "l2_reg": 0.0, "solver": "qr", "tol": 1e-5,
                  "max_iter": 200, "weighted": False, "positive": False, "singular_x_tol": 1e-6}
            for _ in range(2):                                        # warm-up (pinned result pool, allocator)
                res = _harness.call_plugin("pl_lr_pred_f32", inputs, names, kw)
                del res
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            err = f"{type(e).__name__}: {e}"
    if not all_ok(err is None):                                       # collective 1: also the start barrier
        return {"value": None, "unit": "rows/s", "error": err or "the end-to-end leg failed on another rank"}
    k = max(1, args.e2e_steps)
    dt = None
    try:
        t0 = time.perf_counter()
        for _ in range(k):
            res = _harness.call_plugin("pl_lr_pred_f32", inputs, names, kw)
            assert len(res) == rows
            del res
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / k
    except Exception as e:  # noqa: BLE001
        err = f"{type(e).__name__}: {e}"
    if not all_ok(err is None):                                       # collective 2
        return {"value": None, "unit": "rows/s", "error": err or "the end-to-end leg failed on another rank"}
    return {"value": rows / dt, "unit": "rows/s", "h2d_bytes_per_step": (p + 1) * rows * 4,
            "d2h_bytes_per_step": 2 * rows * 4, "ms_per_step": dt * 1e3, "steps": k,
            "api": "_polars_plugin_pl_lr_pred_f32 (Arrow C data, pinned host buffers)"}


if __name__ == "__main__":
    main()
