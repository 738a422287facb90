#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden_vectors.py tests/test_gpu_models.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --config C3 --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_c3.json 2>gpurun_out/bench_c3.err; python -c "import json;d=json.load(open('gpurun_out/bench_c3.json'));print('C3', round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['parity'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_C3.csv python bench.py --config C3 --steps 3 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1; python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_C3.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; kn=h.index('Kernel Name'); mv=h.index('Metric Value')
for r in rows[-9:]: print(r[mv], r[kn][:80])
PY
