#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden_vectors.py tests/test_gpu_models.py tests/test_gpu_shards.py -m gpu -x -q > gpurun_out/pytest_k6.log 2>&1; grep -E "^E |FAILED|passed|failed" gpurun_out/pytest_k6.log | head -10
timeout 300 python bench.py --config C4 --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_c4.json 2>gpurun_out/bench_c4.err; python -c "import json;d=json.load(open('gpurun_out/bench_c4.json'));print('C4', round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['parity'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_C4.csv python bench.py --config C4 --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1; python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_C4.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; kn=h.index('Kernel Name'); mv=h.index('Metric Value')
for r in rows[-3:]: print(r[mv], r[kn][:80])
PY
