#!/bin/bash
# Round-2 state check: full GPU suite, then one bench line per BASELINE config (device-resident + e2e + CPU arm).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/box.txt; nproc >> gpurun_out/box.txt; free -g >> gpurun_out/box.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/pytest_gpu.log
for c in C2 C1 C3 C4 C5; do
  timeout 600 python bench.py --config $c > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "bench $c rc $?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$c.json'))
    print('$c', 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'kms', round(d['roofline']['kernel_ms'],3), 'e2e', d['e2e'] and d['e2e'].get('value'), 'pinned', d['e2e'] and d['e2e'].get('pinned_value'), 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'], 'parity', d['parity'])
except Exception as e:
    print('$c', 'FAILED', e)
PY
done
