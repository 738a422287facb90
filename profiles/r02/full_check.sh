#!/bin/bash
# full GPU suite + one bench line per BASELINE config; K5 staged on/off
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/pytest_gpu.log
for st in 1 0; do PDSB_K5_STAGED=$st timeout 300 python bench.py --config C3 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_c3_st$st.json 2>gpurun_out/bench_c3_st$st.err; python -c "import json;d=json.load(open('gpurun_out/bench_c3_st$st.json'));print('K5 staged=$st', d['ms_per_step'], d['roofline']['frac'], d['parity'])"; done
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
