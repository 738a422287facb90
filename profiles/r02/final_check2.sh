#!/bin/bash
# after the CPU-quota fix: CPU arm at the quota's thread count vs all 128 reported cores; default bench line (pageable e2e with quota-sized staging)
mkdir -p gpurun_out
echo "cpu.max: $(cat /sys/fs/cgroup/cpu.max 2>/dev/null)"
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_reference_quota.json 2> gpurun_out/bench_reference_quota.err; echo "reference (quota threads) rc $?"
PDSB_BENCH_THREADS=128 timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_reference_128.json 2> gpurun_out/bench_reference_128.err; echo "reference (128 threads) rc $?"
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default rc $?"
python - <<'PY'
import json
for f in ("bench_reference_quota", "bench_reference_128", "bench_default"):
    d = json.load(open(f"gpurun_out/{f}.json"))
    e = d.get("e2e") or {}
    cb = d.get("cpu_baseline") or {}
    print(f, "| value %.4g" % d["value"], "| ms", round(d.get("ms_per_step", 0), 3), "| e2e %.4g" % (e.get("value") or 0), e.get("ms_per_step"), "pinned", e.get("pinned_ms_per_step"), "| cpu", cb.get("value"), cb.get("cores"))
PY
